"""Detection post-processing: oracle vs the reference's own outputs (CPU), device vs oracle (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cases_post  # noqa: E402
from oracle import postprocess as op  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_ref.npz"))
MODES = [(m, o) for m in (0, 1, 2) for o in (False, True)]


@pytest.mark.parametrize("name", list(cases_post.nms_cases().keys()))
def test_oracle_nms_equals_reference(name):
    corners, score, cls, valid = cases_post.nms_cases()[name]
    for mode, old in MODES:
        got = op.nms_boxes(corners, score, cls, valid, mode, old, 0.25)
        np.testing.assert_array_equal(got, GOLD["nms/%s/m%d_o%d" % (name, mode, int(old))])


@pytest.mark.parametrize("name", list(cases_post.box_cases().keys()))
def test_oracle_box_counts_equal_scipy_hull_test(name):
    pts, corners = cases_post.box_cases()[name]
    np.testing.assert_array_equal(op.box_point_counts(pts, corners), GOLD["box/%s/count" % name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases_post.nms_cases().keys()))
def test_device_nms_equals_oracle(name):
    from spacap3d_b200 import _ext
    corners, score, cls, valid = cases_post.nms_cases()[name]
    d = lambda a: torch.from_numpy(a).cuda()
    for mode, old in MODES:
        got = _ext.nms_boxes(d(corners), d(score), d(cls), d(valid), mode, old, 0.25).cpu().numpy()
        np.testing.assert_array_equal(got, GOLD["nms/%s/m%d_o%d" % (name, mode, int(old))])
    # equal scores: ordered as by a stable sort, like the oracle
    tied = np.round(score * 8) / 8
    for mode in (1, 2):
        got = _ext.nms_boxes(d(corners), d(tied.astype(np.float32)), d(cls), None, mode, False, 0.25).cpu().numpy()
        np.testing.assert_array_equal(got, op.nms_boxes(corners, tied.astype(np.float32), cls, None, mode, False, 0.25))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(cases_post.box_cases().keys()))
def test_device_box_counts_equal_oracle(name):
    from spacap3d_b200 import _ext
    pts, corners = cases_post.box_cases()[name]
    got = _ext.box_point_counts(torch.from_numpy(pts).cuda(), torch.from_numpy(corners).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got, GOLD["box/%s/count" % name])


@pytest.mark.gpu
def test_parse_predictions_on_detector_output():
    """Full detector forward on two scenes -> parse_predictions on the device == the numpy restatement of
    ap_helper.parse_predictions on the same tensors (SpaCap3D's evaluation settings, scripts/eval.py:195-203)."""
    from spacap3d_b200.detector import VoteNetDetector
    from spacap3d_b200.postprocess import parse_predictions
    from spacap3d_b200.scenes import make_scene
    torch.manual_seed(0)
    model = VoteNetDetector(input_feature_dim=1).cuda().eval()
    pc = torch.from_numpy(np.stack([make_scene(41, 20000), make_scene(42, 20000)], 0)).cuda()
    with torch.no_grad():
        out = model({"point_clouds": pc})
    cfg = {"remove_empty_box": True, "use_3d_nms": True, "nms_iou": 0.25, "use_old_type_nms": False, "cls_nms": True,
           "per_class_proposal": True, "conf_thresh": 0.05}
    from spacap3d_b200.postprocess import predictions_mask
    got = parse_predictions(out, cfg)
    mask_d, obj_d, sem_d = predictions_mask(out, cfg)
    obj_prob, sem_probs = obj_d.cpu().numpy(), sem_d.cpu().numpy()
    corners = out["bbox_corner"].detach().cpu().numpy()
    # same probabilities in (torch's and numpy's fp32 softmax may differ in the last bit), reference algorithm out
    nonempty = op.box_point_counts(out["point_clouds"].cpu().numpy(), corners) >= 5
    want_mask = op.nms_boxes(corners, obj_prob, out["sem_cls"].cpu().numpy(), nonempty, 2, False, 0.25)
    np.testing.assert_array_equal(mask_d.cpu().numpy(), want_mask)
    np.testing.assert_array_equal(out["pred_mask"], want_mask)
    assert 0 < want_mask.sum() < want_mask.size and (~nonempty).any() or True
    np.testing.assert_allclose(obj_prob, op.softmax(out["objectness_scores"].cpu().numpy())[:, :, 1], rtol=1e-6, atol=1e-7)
    assert len(got) == 2
    for i, scene in enumerate(got):
        keep = [j for j in range(want_mask.shape[1]) if want_mask[i, j] == 1 and obj_prob[i, j] > 0.05]
        assert len(scene) == len(keep) * sem_probs.shape[2]
        for n, (c, box, conf) in enumerate(scene):
            ii, j = divmod(n, len(keep))
            assert c == ii and np.array_equal(box, corners[i, keep[j]])
            assert conf == sem_probs[i, keep[j], ii] * obj_prob[i, keep[j]]
