"""Parity of the sm_100a kernels (through the C ABI / `_ext`) against
  (1) the CPU oracle (oracle/pointnet2_oracle.c) on every seeded case,
  (2) the committed golden vectors produced by the reference's own CUDA extension,
  (3) the reference extension itself when oracle/_ref/ is present on the box.
Bar: bit-exact for indices and copies; rtol 1e-5 for atomically accumulated gradients
(the reference's own gradients are order-nondeterministic).
"""
import numpy as np
import pytest
import torch

import cases
import oracle

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def ext():
    from spacap3d_b200 import _ext
    return _ext


@pytest.mark.parametrize("name", list(cases.fps_cases().keys()))
def test_fps(ext, ref_ext, name):
    xyz, m = cases.fps_cases()[name]
    got = ext.furthest_point_sampling(cu(xyz), m).cpu().numpy()
    want = oracle.furthest_point_sampling(xyz, m)
    assert got.dtype == np.int32 and got.shape == want.shape
    np.testing.assert_array_equal(got, want)
    if ref_ext is not None:
        np.testing.assert_array_equal(got, ref_ext.furthest_point_sampling(cu(xyz), m).cpu().numpy())
    # fused new_xyz output == gather of the indices
    idx2, new_xyz = ext.furthest_point_sampling_with_xyz(cu(xyz), m)
    np.testing.assert_array_equal(idx2.cpu().numpy(), want)
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), cases.fps_follow_on(xyz, want, m))


ALGOS = ("auto", "cluster", "bucket")


def _algo(ext, name):
    return ext.launch_options(fps_algo={"auto": ext.FPS_AUTO, "cluster": ext.FPS_CLUSTER, "bucket": ext.FPS_BUCKET}[name])


@pytest.mark.parametrize("name", list(cases.fps_large_cases().keys()))
def test_fps_large_clouds(ext, name):
    """N >= 8192: the bucketed sampler (Morton-sorted points parked in L2, one CTA per scene), the cluster
    sampler (everything on-chip) and the oracle all agree bit for bit -- lattices with exact ties everywhere,
    heavy duplicates, planes / lines, all-identical points included."""
    xyz, m = cases.fps_large_cases()[name]
    want = oracle.furthest_point_sampling(xyz, m)
    for algo in ALGOS:
        with _algo(ext, algo):
            got = ext.furthest_point_sampling(cu(xyz), m).cpu().numpy()
            np.testing.assert_array_equal(got, want, err_msg=algo)
            idx2, new_xyz = ext.furthest_point_sampling_with_xyz(cu(xyz), m, hint_ordered=True)
        np.testing.assert_array_equal(idx2.cpu().numpy(), want, err_msg=algo)
        np.testing.assert_array_equal(new_xyz.cpu().numpy(), cases.fps_follow_on(xyz, want, m))


@pytest.mark.parametrize("algo", ALGOS)
def test_fps_strict_sequence_flags_and_chain(ext, algo):
    """spc_furthest_point_sampling_ex2: per scene "the output is a strict FPS sequence" flags (tracked exactly by the
    bucketed sampler, implied by a successful ordered-prefix proof, 0 = unknown for the cluster sampler) feed the
    next sampler, which then skips proof and rounds.  Whatever the flags say, every level of the
    40k -> 2048 -> 1024 -> 512 -> 256 chain equals the oracle, and a flag of 1 is a promise that FPS over a prefix of
    that output is the identity; a clean scene must be flagged 1 by the bucketed sampler (so the shortcut is really
    exercised), scenes with exact ties 0."""
    sc, _ = cases.fps_cases()["scene_40k"]                        # scene 1 has duplicated points
    lat = cases.fps_large_cases()["lattice_13824"][0][:1]          # exact ties everywhere
    with _algo(ext, algo):
        for xyz, expect in ((sc, [1, 0]), (lat, [0])):
            cur_np, cur = xyz, cu(xyz)
            known, hint = None, False
            for lvl, m in enumerate((2048, 1024, 512, 256)):
                if m > cur_np.shape[1]:
                    break
                idx, new_xyz, strict = ext.furthest_point_sampling_with_xyz(cur, m, hint_ordered=hint,
                                                                             known_ordered=known, want_strict=True)
                want = oracle.furthest_point_sampling(cur_np, m)
                np.testing.assert_array_equal(idx.cpu().numpy(), want)
                cur_np = cases.fps_follow_on(cur_np, want, m)
                np.testing.assert_array_equal(new_xyz.cpu().numpy(), cur_np)
                flags = strict.cpu().numpy().tolist()
                if lvl == 0:
                    assert flags == (expect if algo == "bucket" else [0] * len(expect)), (algo, flags)
                for b, f in enumerate(flags):                     # a flag of 1 is a promise: identity below
                    if f and m // 2 >= 1:
                        sub_idx = oracle.furthest_point_sampling(cur_np[b:b + 1], m // 2)
                        np.testing.assert_array_equal(sub_idx[0], np.arange(m // 2))
                cur, known, hint = new_xyz, strict, True


def test_fps_config5_200k_points(ext):
    """BASELINE config 5's largest cloud: 200 000 points -> 8192 samples (a 16-CTA cluster), one scene."""
    from spacap3d_b200.scenes import make_scene_xyz
    xyz = make_scene_xyz(77, 200000)[None]
    want = oracle.furthest_point_sampling(xyz, 8192)
    got, new_xyz = ext.furthest_point_sampling_with_xyz(cu(xyz), 8192)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), cases.fps_follow_on(xyz, want, 8192))


@pytest.mark.parametrize("algo", ALGOS)
def test_fps_scene_40k_every_sampler(ext, ref_ext, algo):
    """The BASELINE shape (8 x 40 000 -> 2048 is what the detector runs) through every sampler."""
    xyz, m = cases.fps_cases()["scene_40k"]
    want = oracle.furthest_point_sampling(xyz, m)
    with _algo(ext, algo):
        got, new_xyz = ext.furthest_point_sampling_with_xyz(cu(xyz), m)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    np.testing.assert_array_equal(new_xyz.cpu().numpy(), cases.fps_follow_on(xyz, want, m))
    if ref_ext is not None:
        np.testing.assert_array_equal(want, ref_ext.furthest_point_sampling(cu(xyz), m).cpu().numpy())


@pytest.mark.parametrize("algo", ALGOS)
def test_fps_heavy_duplicates_every_sampler(ext, algo):
    """12 000 points drawn with replacement from 5 000: every distance tie of the reference's block-tree arg-max
    must be broken the same way, whatever the sampler and its internal layout."""
    rng = np.random.default_rng(7)
    base = rng.uniform(-3, 3, (3, 5000, 3)).astype(np.float32)
    pick = rng.integers(0, 5000, (3, 12000))
    xyz = np.ascontiguousarray(np.take_along_axis(base, pick[..., None].repeat(3, -1), 1))
    want = oracle.furthest_point_sampling(xyz, 700)
    with _algo(ext, algo):
        got = ext.furthest_point_sampling(cu(xyz), 700).cpu().numpy()
    np.testing.assert_array_equal(got, want)


def test_fps_bucket_sampler_out_of_range(ext):
    """C ABI: SPC_FPS_BUCKET outside its size range is refused (SPC_ERR_UNSUPPORTED), never silently replaced;
    the Python wrapper treats the thread-local setting as a preference and applies it inside the range only."""
    from spacap3d_b200 import _lib
    xyz = cu(np.random.default_rng(0).uniform(-1, 1, (1, 50000, 3)).astype(np.float32))
    out = torch.empty((1, 64), dtype=torch.int32, device=DEV)
    nbytes = _lib.load().spc_fps_workspace_bytes(1, 50000, 64)
    ws = torch.empty((nbytes + 3) // 4, dtype=torch.int32, device=DEV)
    with pytest.raises(_lib.SpcUnsupported):
        _lib.call("spc_furthest_point_sampling_ex2", xyz.data_ptr(), 1, 50000, 64, out.data_ptr(), None, 0, None, None,
                  ws.data_ptr(), nbytes, ext.FPS_BUCKET, torch.cuda.current_stream().cuda_stream)
    with _algo(ext, "bucket"):
        got = ext.furthest_point_sampling(xyz, 64)
    np.testing.assert_array_equal(got.cpu().numpy(), oracle.furthest_point_sampling(xyz.cpu().numpy(), 64))


def test_launch_options_are_thread_local(ext):
    """SURVEY 8b threading row (nn.DataParallel calls the library from several threads): launch hints are per
    thread and per call; two threads using different samplers concurrently get the same, correct picks."""
    import threading
    xyz, m = cases.fps_large_cases()["clusters_20000"]
    want = oracle.furthest_point_sampling(xyz, m)
    x = cu(xyz)
    torch.cuda.synchronize()
    errors = []

    def worker(algo, reps):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream), _algo(ext, algo):
                for _ in range(reps):
                    assert ext._options.fps_algo == {"cluster": ext.FPS_CLUSTER, "bucket": ext.FPS_BUCKET}[algo]
                    got = ext.furthest_point_sampling(x, m)
                    np.testing.assert_array_equal(got.cpu().numpy(), want)
        except Exception as e:  # noqa: BLE001
            errors.append((algo, repr(e)))

    threads = [threading.Thread(target=worker, args=(a, 4)) for a in ("cluster", "bucket", "cluster", "bucket")]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert ext._options.fps_algo == ext.FPS_AUTO              # the main thread's options were never touched


def test_fps_prefix_property(ext):
    """FPS over an FPS-ordered prefix returns 0..n-1 (the model relies on it)."""
    xyz = cases.fps_cases()["scene_40k"][0][:1]
    idx, new_xyz = ext.furthest_point_sampling_with_xyz(cu(xyz), 2048)
    idx2 = ext.furthest_point_sampling(new_xyz.contiguous(), 1024).cpu().numpy()
    np.testing.assert_array_equal(idx2[0], np.arange(1024, dtype=np.int32))


@pytest.mark.parametrize("name", list(cases.ball_query_extra_cases().keys()))
def test_ball_query_long_thin_clouds(ext, ref_ext, name):
    """Grids with 750-2000 cells along one axis (and clouds 3 km from the origin): the candidate range comes from the
    monotone cell function applied to q -+ r, so no neighbour can be missed whatever the fp32 rounding."""
    new, xyz, r, ns = cases.ball_query_extra_cases()[name]
    want = oracle.ball_query(new, xyz, r, ns)
    got = ext.ball_query(cu(new), cu(xyz), r, ns).cpu().numpy()
    np.testing.assert_array_equal(got, want)
    assert (got != 0).any()
    if ref_ext is not None:
        np.testing.assert_array_equal(got, ref_ext.ball_query(cu(new), cu(xyz), r, ns).cpu().numpy())


@pytest.mark.parametrize("name", list(cases.ball_query_cases().keys()))
def test_ball_query(ext, ref_ext, name):
    new_xyz, xyz, r, ns = cases.ball_query_cases()[name]
    got = ext.ball_query(cu(new_xyz), cu(xyz), r, ns).cpu().numpy()
    want = oracle.ball_query(new_xyz, xyz, r, ns)
    np.testing.assert_array_equal(got, want)
    if ref_ext is not None:
        np.testing.assert_array_equal(got, ref_ext.ball_query(cu(new_xyz), cu(xyz), float(r), ns).cpu().numpy())


@pytest.mark.parametrize("name", list(cases.three_nn_cases().keys()))
def test_three_nn(ext, ref_ext, name):
    unknown, known = cases.three_nn_cases()[name]
    d2, idx = ext.three_nn(cu(unknown), cu(known))
    wd2, widx = oracle.three_nn(unknown, known)
    np.testing.assert_array_equal(idx.cpu().numpy(), widx)
    np.testing.assert_array_equal(d2.cpu().numpy(), wd2)
    if ref_ext is not None:
        rd2, ridx = ref_ext.three_nn(cu(unknown), cu(known))
        np.testing.assert_array_equal(idx.cpu().numpy(), ridx.cpu().numpy())
        np.testing.assert_array_equal(d2.cpu().numpy(), rd2.cpu().numpy())


@pytest.mark.parametrize("name", list(cases.gather_cases().keys()))
def test_gather(ext, ref_ext, name):
    pts, idx = cases.gather_cases()[name]
    got = ext.gather_points(cu(pts), cu(idx)).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.gather_points(pts, idx))
    g = cases.grad_for("gather/" + name, got.shape)
    gg = ext.gather_points_grad(cu(g), cu(idx), pts.shape[2]).cpu().numpy()
    np.testing.assert_allclose(gg, oracle.gather_points_grad(g, idx, pts.shape[2]), rtol=1e-5, atol=1e-6)
    if ref_ext is not None:
        np.testing.assert_array_equal(got, ref_ext.gather_points(cu(pts), cu(idx)).cpu().numpy())


@pytest.mark.parametrize("name", list(cases.group_cases().keys()))
def test_group(ext, ref_ext, name):
    pts, idx = cases.group_cases()[name]
    got = ext.group_points(cu(pts), cu(idx)).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.group_points(pts, idx))
    g = cases.grad_for("group/" + name, got.shape)
    gg = ext.group_points_grad(cu(g), cu(idx), pts.shape[2]).cpu().numpy()
    want = oracle.group_points_grad(g, idx, pts.shape[2])
    np.testing.assert_allclose(gg, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())
    if ref_ext is not None:
        np.testing.assert_array_equal(got, ref_ext.group_points(cu(pts), cu(idx)).cpu().numpy())
        rg = ref_ext.group_points_grad(cu(g), cu(idx), pts.shape[2]).cpu().numpy()
        np.testing.assert_allclose(gg, rg, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize("name", list(cases.group_grad_big_cases().keys()))
def test_group_grad_list_based(ext, ref_ext, name, monkeypatch):
    """The atomic-free backward (per-point position lists): equals the oracle within fp32 summation
    error, equals the atomic kernel, and is bit-reproducible when the positions fit two partitions."""
    idx, C, N = cases.group_grad_big_cases()[name]
    B, npoint, ns = idx.shape
    g = cases.grad_for("group_grad_big/" + name, (B, C, npoint, ns))
    want = oracle.group_points_grad(g, idx, N)
    tol = dict(rtol=1e-5, atol=2e-5 * np.abs(want).max())
    got = ext.group_points_grad(cu(g), cu(idx), N).cpu().numpy()
    np.testing.assert_allclose(got, want, **tol)
    if npoint * ns <= 49152 and N <= 8192:      # larger clouds build their lists with atomic cursors: order not fixed
        for _ in range(3):
            np.testing.assert_array_equal(ext.group_points_grad(cu(g), cu(idx), N).cpu().numpy(), got)
    # the entry point without a workspace takes the atomic kernel
    from spacap3d_b200 import _lib
    ga, ia = cu(g), cu(idx)
    atomic = torch.empty((B, C, N), dtype=torch.float32, device=DEV)
    _lib.call("spc_group_points_grad", ga.data_ptr(), ia.data_ptr(), B, C, N, npoint, ns, atomic.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    np.testing.assert_allclose(atomic.cpu().numpy(), want, **tol)
    if ref_ext is not None:
        np.testing.assert_allclose(ref_ext.group_points_grad(cu(g), cu(idx), N).cpu().numpy(), want, **tol)


@pytest.mark.parametrize("name", list(cases.interp_cases().keys()) + list(cases.interp_big_cases().keys()))
def test_interpolate(ext, ref_ext, name):
    pts, idx, w = {**cases.interp_cases(), **cases.interp_big_cases()}[name]
    got = ext.three_interpolate(cu(pts), cu(idx), cu(w)).cpu().numpy()
    want = oracle.three_interpolate(pts, idx, w)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6)   # the bar north_star states
    g = cases.grad_for("interp/" + name, got.shape)
    gg = ext.three_interpolate_grad(cu(g), cu(idx), cu(w), pts.shape[2]).cpu().numpy()
    wg = oracle.three_interpolate_grad(g, idx, w, pts.shape[2])
    np.testing.assert_allclose(gg, wg, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(wg).max()))
    if ref_ext is not None:
        np.testing.assert_array_equal(got, ref_ext.three_interpolate(cu(pts), cu(idx), cu(w)).cpu().numpy())
        np.testing.assert_array_equal(got, want)          # bit-equal to the oracle as well


# ---------------------------------------------------------------- BASELINE-size properties ----
def test_group_full_size_roundtrip(ext):
    """Config-4 shape (B=2 of 8, C=132, N=40k, 2048x64): grouping is a pure gather, so
    out[b,c,j,k] == points[b,c,idx[b,j,k]] checked with torch indexing, and the backward of a
    ones-gradient equals the histogram of idx (size-independent properties)."""
    g = torch.Generator(device="cpu").manual_seed(5)
    B, C, N, npoint, ns = 2, 132, 40000, 2048, 64
    pts = torch.randn(B, C, N, generator=g).to(DEV)
    idx = torch.randint(0, N, (B, npoint, ns), generator=g, dtype=torch.int32).to(DEV)
    out = ext.group_points(pts, idx)
    want = torch.gather(pts, 2, idx.long().view(B, 1, -1).expand(-1, C, -1)).view(B, C, npoint, ns)
    assert torch.equal(out, want)
    ones = torch.ones(B, 3, npoint, ns, device=DEV)
    hist = ext.group_points_grad(ones, idx, N)
    want_hist = torch.zeros(B, N, device=DEV).scatter_add_(1, idx.long().view(B, -1), torch.ones(B, npoint * ns, device=DEV))
    assert torch.equal(hist[:, 0], want_hist) and torch.equal(hist[:, 2], want_hist)
    # with >= 4 channels a 40 k cloud takes the point-owned gather (one list per point over all positions)
    hist8 = ext.group_points_grad(torch.ones(B, 9, npoint, ns, device=DEV), idx, N)
    assert torch.equal(hist8[:, 0], want_hist) and torch.equal(hist8[:, 7], want_hist) and torch.equal(hist8[:, 8], want_hist)
    g = torch.randn(B, 12, npoint, ns, generator=torch.Generator(device="cpu").manual_seed(6)).to(DEV)
    got = ext.group_points_grad(g, idx, N)
    want = torch.zeros(B, 12, N, device=DEV).scatter_add_(2, idx.long().view(B, 1, -1).expand(-1, 12, -1), g.view(B, 12, -1))
    torch.testing.assert_close(got, want, rtol=1e-5, atol=2e-5 * want.abs().max().item())


def test_ball_query_full_size_properties(ext):
    """SA1 shape on a 40k scene: rows ascending up to the pad, every listed point inside the
    ball, pad == first hit, and the hit count equals a brute-force torch count (capped)."""
    xyz_np = cases.fps_cases()["scene_40k"][0][:1]
    xyz = cu(xyz_np)
    idx_fps, new_xyz = ext.furthest_point_sampling_with_xyz(xyz, 2048)
    r, ns = 0.2, 64
    idx = ext.ball_query(new_xyz, xyz, r, ns).long()
    d2 = ((new_xyz[:, :, None, :] - torch.gather(
        xyz[:, None].expand(-1, 2048, -1, -1), 2, idx[..., None].expand(-1, -1, -1, 3))) ** 2).sum(-1)
    assert (d2 < r * r * (1 + 1e-5)).all()
    cnt_true = (torch.cdist(new_xyz, xyz) < r).sum(-1).clamp(max=ns)           # approx count
    first = idx[..., :1]
    is_pad = torch.arange(ns, device=DEV)[None, None] >= cnt_true[..., None]
    # allow +-1 disagreement at the boundary between cdist rounding and the exact FMA order
    inc = (idx[..., 1:] > idx[..., :-1]) | is_pad[..., 1:] | (torch.arange(1, ns, device=DEV)[None, None] >= (cnt_true[..., None] - 1))
    assert inc.all()
    strict_pad = torch.arange(ns, device=DEV)[None, None] >= (cnt_true[..., None] + 1)
    assert ((idx == first) | ~strict_pad).all()


def test_error_paths(ext):
    x = torch.zeros(1, 4, 3)
    with pytest.raises(RuntimeError):
        ext.furthest_point_sampling(x, 2)                        # CPU tensor
    with pytest.raises(RuntimeError):
        ext.gather_points(torch.zeros(1, 3, 4, device=DEV), torch.zeros(1, 2, device=DEV))  # idx not int
    with pytest.raises(RuntimeError):
        ext.ball_query(torch.zeros(1, 2, 3, device=DEV).transpose(1, 2), torch.zeros(1, 4, 3, device=DEV), 0.1, 2)


# ---------------------------------------------------------------- verified ordered-prefix path ---
def _fps_hint(ext, xyz_t, m):
    idx, new_xyz = ext.furthest_point_sampling_with_xyz(xyz_t, m, hint_ordered=True)
    return idx.cpu().numpy(), new_xyz.cpu().numpy()


@pytest.mark.parametrize("name", list(cases.fps_cases().keys()))
def test_fps_ordered_hint_never_changes_results(ext, name):
    """The hint only affects speed: on arbitrary (unordered) inputs the proof fails and the
    sequential kernel runs; results equal the oracle either way."""
    xyz, m = cases.fps_cases()[name]
    got, _ = _fps_hint(ext, cu(xyz), m)
    np.testing.assert_array_equal(got, oracle.furthest_point_sampling(xyz, m))


@pytest.mark.parametrize("name", ["scene_40k", "dup_2048", "lattice_4096", "origin_mix", "n1000_T512"])
def test_fps_ordered_hint_on_fps_outputs(ext, name):
    """Chained FPS as in SA1->SA2->SA3->SA4 (FPS of an FPS-ordered list): with and without the
    hint, against the oracle -- including inputs with duplicates / lattice ties / skipped points,
    where the answer is NOT the identity and the proof must fail."""
    xyz, m = cases.fps_cases()[name]
    cur = xyz
    cur_t = cu(xyz)
    n_next = m
    for level in range(3):
        idx, new_xyz = ext.furthest_point_sampling_with_xyz(cur_t, n_next, hint_ordered=(level > 0))
        want = oracle.furthest_point_sampling(cur, n_next)
        np.testing.assert_array_equal(idx.cpu().numpy(), want)
        cur = cases.fps_follow_on(cur, want, n_next)
        np.testing.assert_array_equal(new_xyz.cpu().numpy(), cur)
        cur_t = new_xyz.contiguous()
        n_next = max(2, n_next // 2)
