"""parse_predictions with the heavy parts on the device.

Same call and result as the reference's lib/ap_helper.py:44-160 -- `parse_predictions(end_points, config_dict)`
returns `batch_pred_map_cls` (per scene a list of `(class, corners (8,3) float64 ndarray, confidence)`) and stores
`end_points['pred_mask']` and `end_points['batch_pred_map_cls']` -- but the B*K scipy hull tests over 40 k points
(ap_helper.py:69-79) and the per-scene numpy NMS (:82-137, utils/nms.py) run as two CUDA kernels on the tensors the
detector already holds on the GPU; only the final list assembly (needs Python objects) touches the host, after one
device->host copy of (corners, probabilities, mask).
"""
import numpy as np
import torch

from . import _ext


def predictions_mask(end_points, config_dict):
    """Device part: returns (pred_mask (B,K) int32, obj_prob (B,K) f32, sem_cls_probs (B,K,C) f32), all on the GPU."""
    corners = end_points["bbox_corner"]
    if not torch.is_tensor(corners):
        raise TypeError("bbox_corner must be a device tensor (detector.decode_pred_box keeps it on the GPU)")
    corners = corners.detach().to(torch.float64).contiguous()
    sem_cls_probs = torch.softmax(end_points["sem_cls_scores"].detach().float(), dim=-1)
    obj_prob = torch.softmax(end_points["objectness_scores"].detach().float(), dim=-1)[:, :, 1].contiguous()
    valid = None
    if config_dict["remove_empty_box"]:
        pts = end_points["point_clouds"].detach().float().contiguous()
        valid = (_ext.box_point_counts(pts, corners) >= 5).to(torch.int32)
    if not config_dict["use_3d_nms"]:
        mode = 0
    elif not config_dict["cls_nms"]:
        mode = 1
    else:
        mode = 2
    cls = end_points["sem_cls"].detach().to(torch.int64).contiguous() if mode == 2 else None
    pred_mask = _ext.nms_boxes(corners, obj_prob, cls, valid, mode, bool(config_dict["use_old_type_nms"]),
                               float(config_dict["nms_iou"]))
    return pred_mask, obj_prob, sem_cls_probs


def parse_predictions(end_points, config_dict):
    pred_mask_d, obj_prob_d, sem_probs_d = predictions_mask(end_points, config_dict)
    corners = end_points["bbox_corner"].detach().cpu().numpy()
    pred_mask = pred_mask_d.cpu().numpy()
    obj_prob = obj_prob_d.cpu().numpy()
    sem_cls_probs = sem_probs_d.cpu().numpy()
    pred_sem_cls = end_points["sem_cls"]
    end_points["pred_mask"] = pred_mask.astype(np.float64)      # the reference stores a float64 0/1 array
    bsize, K = pred_mask.shape
    batch_pred_map_cls = []
    for i in range(bsize):
        keep = [j for j in range(K) if pred_mask[i, j] == 1 and obj_prob[i, j] > config_dict["conf_thresh"]]
        if config_dict["per_class_proposal"]:
            cur_list = []
            num_class = (config_dict["dataset_config"].num_class if "dataset_config" in config_dict
                         else sem_cls_probs.shape[2])
            for ii in range(num_class):
                cur_list += [(ii, corners[i, j], sem_cls_probs[i, j, ii] * obj_prob[i, j]) for j in keep]
            batch_pred_map_cls.append(cur_list)
        else:
            batch_pred_map_cls.append([(pred_sem_cls[i, j].item(), corners[i, j], obj_prob[i, j]) for j in keep])
    end_points["batch_pred_map_cls"] = batch_pred_map_cls
    return batch_pred_map_cls
