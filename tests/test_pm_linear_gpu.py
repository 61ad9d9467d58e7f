"""spc_pm_linear (tcgen05 point-major 1x1-conv layer, fp16-pair operands, three MMAs per product) against a float64
torch evaluation of the same layer.  The bar is fp32-grade: element-wise |err| <= 2e-5 (|ref| + rms(ref)) -- two
orders of magnitude inside north_star's 1e-2 for the 16-bit MLP path, so that the FP / voting / proposal layers add
nothing measurable to the fused set-abstraction kernels' error."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 2e-5


def _nerr(x, ref):
    x, ref = x.double(), ref.double()
    rms = ref.pow(2).mean().sqrt().clamp_min(1e-30)
    return ((x - ref).abs() / (ref.abs() + rms)).max().item()


def _layer(M, K, N, seed, pair_input):
    from spacap3d_b200 import _ext
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = (torch.randn(N, generator=g) * 0.1).to(DEV)
    if pair_input:
        X = _ext.split_half(x)
        x_eff = X[0].double() + X[1].double()
    else:
        X = x.to(_ext.HALF)
        x_eff = X.double()
    ref = x_eff @ W.double().t() + b.double()
    return X, _ext.split_half(W), b, ref


@pytest.mark.parametrize("M,K,N,pair", [(4096, 512, 256, False), (8192, 256, 256, True), (2048, 128, 128, True),
                                        (300, 64, 32, False), (128 * 3 + 7, 192, 96, True), (1000, 136, 64, False),
                                        (257, 8, 32, True)])
def test_hidden_layer(M, K, N, pair):
    from spacap3d_b200 import _ext
    X, W, b, ref = _layer(M, K, N, 1, pair)
    hi, lo = _ext.pm_linear(X, W, b, _ext.PM_HIDDEN, M)
    got = hi.double() + lo.double()
    want = torch.relu(ref)
    assert _nerr(got, want) <= TOL
    assert _nerr(hi, want) <= 1e-3                      # the hi half alone is the fp16 rounding of the result


@pytest.mark.parametrize("B,n,K,N", [(8, 1024, 512, 256), (2, 512, 256, 256), (3, 256, 128, 128)])
def test_channel_major_output(B, n, K, N):
    from spacap3d_b200 import _ext
    X, W, b, ref = _layer(B * n, K, N, 2, True)
    out, (hi, lo) = _ext.pm_linear(X, W, b, _ext.PM_OUT_CM, n)
    want = torch.relu(ref).view(B, n, N).transpose(1, 2)
    assert out.shape == (B, N, n) and out.dtype == torch.float32
    assert _nerr(out, want) <= TOL
    assert _nerr(hi.double() + lo.double(), torch.relu(ref)) <= TOL


@pytest.mark.parametrize("B,n,K,N", [(8, 256, 128, 79), (1, 128, 64, 5), (2, 256, 128, 272)])
def test_point_major_f32_output_any_width(B, n, K, N):
    from spacap3d_b200 import _ext
    X, W, b, ref = _layer(B * n, K, N, 3, True)
    out = _ext.pm_linear(X, W, b, _ext.PM_OUT_PM32, n)
    assert out.shape == (B * n, N)
    assert _nerr(out, ref) <= TOL                        # no activation (proposal head's last conv)


@pytest.mark.parametrize("B,n,D", [(8, 1024, 256), (2, 128, 64)])
def test_vote_tail(B, n, D):
    """conv3 + offsets + residual + L2 normalisation (models/voting_module.py:52-61, models/SpaCapNet.py:66-67)."""
    from spacap3d_b200 import _ext
    X, W, b, ref = _layer(B * n, D, 3 + D, 4, True)
    g = torch.Generator(device="cpu").manual_seed(9)
    seed_xyz = torch.randn(B, n, 3, generator=g).to(DEV)
    seed_cm = torch.relu(torch.randn(B, D, n, generator=g)).to(DEV)
    vote_xyz, out, (hi, lo) = _ext.pm_linear(X, W, b, _ext.PM_VOTE, n, seed_cm=seed_cm, seed_xyz=seed_xyz)
    net = ref.view(B, n, 3 + D)                         # feature rows first, the three xyz-offset rows last
    want_xyz = seed_xyz.double() + net[..., D:]
    v = seed_cm.double().transpose(1, 2) + net[..., :D]
    v = v / v.norm(dim=2, keepdim=True)
    assert _nerr(vote_xyz, want_xyz) <= TOL
    assert _nerr(out, v.transpose(1, 2)) <= TOL
    assert _nerr((hi.double() + lo.double()).view(B, n, D), v) <= TOL


def test_argument_checks():
    from spacap3d_b200 import _ext, _lib
    X, W, b, _ = _layer(256, 64, 32, 5, False)
    with pytest.raises(_lib.SpcError):
        _ext.pm_linear(X[:, :44].contiguous(), (W[0][:, :44].contiguous(), W[1][:, :44].contiguous()), b,
                       _ext.PM_HIDDEN, 256)          # K not a multiple of 8


def test_launch_hints_do_not_change_results():
    """The per-call launch hints only change the grid and the schedule: output channels per CTA (one CTA per row tile,
    two 128-channel slices = default, four 64-channel slices) and row tiles per CTA (1 = default; more: the CTA loops
    with a double-buffered accumulator) give the same bits -- also with a ragged last tile and in the vote tail,
    whose 272-column accumulator is single-buffered."""
    from spacap3d_b200 import _ext
    M = 128 * 5 + 7                                               # six row tiles, the last one ragged
    X, W, b, _ = _layer(M, 256, 256, 21, True)
    base = None
    for nt, tpc in ((0, 0), (64, 0), (256, 0), (0, 2), (256, 4), (256, 16), (128, 5), (64, 3)):   # (0,2) (128,5) (64,3): W resident
        with _ext.launch_options(pm_n_tile=nt, pm_tiles_per_cta=tpc):
            y_hi, y_lo = _ext.pm_linear(X, W, b, _ext.PM_HIDDEN, M)
            out, _ = _ext.pm_linear(X, W, b, _ext.PM_OUT_CM, M)
        got = (y_hi.clone(), y_lo.clone(), out.clone())
        if base is None:
            base = got
        assert all(torch.equal(a, c) for a, c in zip(got, base)), (nt, tpc)
    B, n, D = 2, 640, 256
    Xv, Wv, bv, _ = _layer(B * n, D, 3 + D, 4, True)
    g = torch.Generator(device="cpu").manual_seed(9)
    seed_xyz = torch.randn(B, n, 3, generator=g).to(DEV)
    seed_cm = torch.relu(torch.randn(B, D, n, generator=g)).to(DEV)
    ref = None
    for tpc in (0, 3):
        with _ext.launch_options(pm_tiles_per_cta=tpc):
            vote_xyz, out, (hi, lo) = _ext.pm_linear(Xv, Wv, bv, _ext.PM_VOTE, n, seed_cm=seed_cm, seed_xyz=seed_xyz)
        got = (vote_xyz.clone(), out.clone(), hi.clone(), lo.clone())
        if ref is None:
            ref = got
        assert all(torch.equal(a, c) for a, c in zip(got, ref)), tpc
    with pytest.raises(RuntimeError):
        with _ext.launch_options(pm_n_tile=100):                  # not a multiple of 16
            _ext.pm_linear(X, W, b, _ext.PM_HIDDEN, M)
