"""CPU oracle of the detection post-processing that follows the CoAlign forward:
anchor generation, box decoding, direction fix, corner projection, size / z filters, rotated NMS, range mask.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, operation by operation on torch-CPU float32 (same dtypes and operation order as the reference):
  * VoxelPostprocessor.generate_anchor_box   /root/reference/opencood/data_utils/post_processor/voxel_postprocessor.py:30-82
  * VoxelPostprocessor.post_process          voxel_postprocessor.py:243-402   (intermediate fusion: data_dict = {'ego': ...})
  * VoxelPostprocessor.delta_to_boxes3d      voxel_postprocessor.py:404-449
  * limit_period / rotate_points_along_z     /root/reference/opencood/utils/common_utils.py:70-79,105-127
  * boxes_to_corners_3d / project_box3d      /root/reference/opencood/utils/box_utils.py:152-204,278-316
  * remove_large_pred_bbx / remove_bbx_abnormal_z   box_utils.py:840-890 (including the reference's quirk: z_len is
    computed from the y coordinates and only tested for truthiness)
  * nms_rotated                              box_utils.py:693-738
  * mask_boxes_outside_range_numpy           box_utils.py:384-421

Pinning: everything except the polygon IoU is PINNED against the unmodified reference run in the build container
(tests/golden/gen_golden_post.py -> tests/golden/post_*.npz).  The polygon IoU inside nms_rotated is shapely's
(third-party, absent): oracle/rotated_iou.c restates it - PARITY UNPINNED for that function (see its header); the
golden generator feeds the same restatement to the reference code in place of shapely.
Tie order of equal scores in `scores.argsort()[::-1]` (numpy quicksort, unspecified) is restated as "lower candidate
index first"; the fixtures contain no ties.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librotated_iou_oracle.so")


def build_c(force=False):
    src = os.path.join(_HERE, "rotated_iou.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src, "-lm"])
    return _SO


_lib = None


def _iou_lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_c())
        _lib.oracle_quad_iou_one_to_many.restype = None
        _lib.oracle_convex_intersection_area.restype = ctypes.c_double
    return _lib


def quad_iou_one_to_many(box, boxes):
    """IoU of quadrilateral `box` (4,2) with each of `boxes` (n,4,2), float32 like common_utils.compute_iou."""
    lib = _iou_lib()
    box = np.ascontiguousarray(box, dtype=np.float64)
    boxes = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 4, 2)
    out = np.empty(boxes.shape[0], dtype=np.float32)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.oracle_quad_iou_one_to_many(box.ctypes.data_as(dp), boxes.ctypes.data_as(dp), int(boxes.shape[0]),
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return out


def convex_intersection_area(a, b):
    lib = _iou_lib()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    return float(lib.oracle_convex_intersection_area(a.ctypes.data_as(dp), int(a.shape[0]), b.ctypes.data_as(dp),
                                                     int(b.shape[0])))


# ---------------------------------------------------------------------------------------------- anchors
def generate_anchor_box(params):
    """voxel_postprocessor.py:30-82 (float64 numpy, order 'hwl' or 'lhw')."""
    aa = params["anchor_args"]
    W, H = aa["W"], aa["H"]
    r = [math.radians(e) for e in aa["r"]]
    num = aa["num"]
    assert num == len(r)
    vh, vw = aa["vh"], aa["vw"]
    xr = [aa["cav_lidar_range"][0], aa["cav_lidar_range"][3]]
    yr = [aa["cav_lidar_range"][1], aa["cav_lidar_range"][4]]
    fs = aa.get("feature_stride", 2)
    x = np.linspace(xr[0] + vw, xr[1] - vw, W // fs)
    y = np.linspace(yr[0] + vh, yr[1] - vh, H // fs)
    cx, cy = np.meshgrid(x, y)
    cx = np.tile(cx[..., np.newaxis], num)
    cy = np.tile(cy[..., np.newaxis], num)
    cz = np.ones_like(cx) * -1.0
    w = np.ones_like(cx) * aa["w"]
    l = np.ones_like(cx) * aa["l"]
    h = np.ones_like(cx) * aa["h"]
    r_ = np.ones_like(cx)
    for i in range(num):
        r_[..., i] = r[i]
    if params["order"] == "hwl":
        return np.stack([cx, cy, cz, h, w, l, r_], axis=-1)
    if params["order"] == "lhw":
        return np.stack([cx, cy, cz, l, h, w, r_], axis=-1)
    raise ValueError("Unknown bbx order.")


# ---------------------------------------------------------------------------------------------- pieces
def limit_period(val, offset=0.5, period=2 * np.pi):
    return val - torch.floor(val / period + offset) * period            # common_utils.py:78


def delta_to_boxes3d(deltas, anchors):
    """voxel_postprocessor.py:404-449: deltas (N,14,H,W) -> (N, H*W*2, 7)."""
    N = deltas.shape[0]
    deltas = deltas.permute(0, 2, 3, 1).contiguous().view(N, -1, 7)
    boxes3d = torch.zeros_like(deltas)
    a = anchors.view(-1, 7).float()
    a_d = torch.sqrt(a[:, 4] ** 2 + a[:, 5] ** 2)
    a_d = a_d.repeat(N, 2, 1).transpose(1, 2)
    a = a.repeat(N, 1, 1)
    boxes3d[..., [0, 1]] = torch.mul(deltas[..., [0, 1]], a_d) + a[..., [0, 1]]
    boxes3d[..., [2]] = torch.mul(deltas[..., [2]], a[..., [3]]) + a[..., [2]]
    boxes3d[..., [3, 4, 5]] = torch.exp(deltas[..., [3, 4, 5]]) * a[..., [3, 4, 5]]
    boxes3d[..., 6] = deltas[..., 6] + a[..., 6]
    return boxes3d


def boxes_to_corners_3d(boxes3d, order):
    """box_utils.py:152-204."""
    b = boxes3d
    if order == "hwl":
        b = boxes3d[:, [0, 1, 2, 5, 4, 3, 6]]
    template = b.new_tensor(([1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1],
                             [1, -1, 1], [1, 1, 1], [-1, 1, 1], [-1, -1, 1])) / 2
    corners = b[:, None, 3:6].repeat(1, 8, 1) * template[None, :, :]
    ang = b[:, 6]
    cosa, sina = torch.cos(ang), torch.sin(ang)
    zeros, ones = ang.new_zeros(corners.shape[0]), ang.new_ones(corners.shape[0])
    rot = torch.stack((cosa, sina, zeros, -sina, cosa, zeros, zeros, zeros, ones), dim=1).view(-1, 3, 3).float()
    corners = torch.matmul(corners.view(-1, 8, 3)[:, :, 0:3].float(), rot).view(-1, 8, 3)   # common_utils.py:105-127
    corners = corners + b[:, None, 0:3]
    return corners


def project_box3d(box3d, tfm):
    """box_utils.py:278-316."""
    c = box3d.transpose(1, 2)
    ones = torch.ones((c.shape[0], 1, 8))
    c = torch.cat((c, ones), dim=1)
    p = torch.matmul(tfm, c)
    return p[:, :3, :].transpose(1, 2)


def remove_large_pred_bbx(b):
    x_len = torch.max(b[:, :, 0], dim=1)[0] - torch.min(b[:, :, 0], dim=1)[0]
    y_len = torch.max(b[:, :, 1], dim=1)[0] - torch.min(b[:, :, 1], dim=1)[0]
    z_len = torch.max(b[:, :, 1], dim=1)[0] - torch.min(b[:, :, 1], dim=1)[0]     # sic: y again (box_utils.py:863-865)
    index = torch.logical_and(x_len <= 6, y_len <= 6)
    return torch.logical_and(index, z_len)


def remove_bbx_abnormal_z(b):
    zmin = torch.min(b[:, :, 2], dim=1)[0]
    zmax = torch.max(b[:, :, 2], dim=1)[0]
    return torch.logical_and(zmin >= -3, zmax <= 1)


def nms_rotated(boxes, scores, threshold, top=1000):
    """box_utils.py:693-738 with shapely's IoU restated by oracle/rotated_iou.c."""
    if boxes.shape[0] == 0:
        return np.array([], dtype=np.int32)
    boxes = boxes.cpu().detach().numpy()
    scores = scores.cpu().detach().numpy()
    polys = boxes[:, :4, :2]                                            # common_utils.convert_format
    ixs = np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))[:top]   # score desc, index asc on ties
    pick = []
    while len(ixs) > 0:
        i = ixs[0]
        pick.append(i)
        iou = quad_iou_one_to_many(polys[i], polys[ixs[1:]])
        remove = np.where(iou > threshold)[0] + 1
        ixs = np.delete(ixs, remove)
        ixs = np.delete(ixs, 0)
    return np.array(pick, dtype=np.int32)


# ---------------------------------------------------------------------------------------------- the whole thing
def post_process(params, anchor_box, tfm, cls_preds, reg_preds, dir_preds=None):
    """voxel_postprocessor.py:243-402 for data_dict = {'ego': {'transformation_matrix': tfm, 'anchor_box': anchor_box}},
    output_dict = {'ego': {'cls_preds', 'reg_preds', 'dir_preds'}}.  Returns (pred_box3d_tensor (K,8,3), scores (K,))
    or (None, None)."""
    prob = torch.sigmoid(cls_preds.permute(0, 2, 3, 1)).reshape(1, -1)
    batch_box3d = delta_to_boxes3d(reg_preds, anchor_box)
    mask = torch.gt(prob, params["target_args"]["score_threshold"]).view(1, -1)
    assert batch_box3d.shape[0] == 1
    boxes3d = batch_box3d[0][mask[0]].view(-1, 7)
    scores = prob[0][mask[0]]
    if dir_preds is not None and len(boxes3d) != 0:
        dir_offset = params["dir_args"]["dir_offset"]
        num_bins = params["dir_args"]["num_bins"]
        dcp = dir_preds.permute(0, 2, 3, 1).contiguous().reshape(1, -1, num_bins)[mask]
        dir_labels = torch.max(dcp, dim=-1)[1]
        period = 2 * np.pi / num_bins
        dir_rot = limit_period(boxes3d[..., 6] - dir_offset, 0, period)
        boxes3d[..., 6] = dir_rot + dir_offset + period * dir_labels.to(dcp.dtype)
        boxes3d[..., 6] = limit_period(boxes3d[..., 6], 0.5, 2 * np.pi)
    if len(boxes3d) == 0:
        return None, None
    corners = boxes_to_corners_3d(boxes3d, order=params["order"])
    proj = project_box3d(corners, tfm)
    keep = torch.logical_and(remove_large_pred_bbx(proj), remove_bbx_abnormal_z(proj))
    proj = proj[keep]
    scores = scores[keep]
    keep_index = nms_rotated(proj, scores, params["nms_thresh"])
    proj = proj[keep_index]
    scores = scores[keep_index]
    pb = proj.cpu().numpy()
    lr = np.asarray(params["gt_range"], dtype=np.float64)
    m = ((pb >= lr[0:3]) & (pb <= lr[3:6])).all(axis=2)
    m = m.sum(axis=1) >= 8
    return torch.from_numpy(pb[m]), scores[m]


def post_process_stage1(params, anchor_box, cls_preds, reg_preds, unc_preds, dir_preds=None):
    """UncertaintyVoxelPostprocessor.post_process_stage1 (uncertainty_voxel_postprocessor.py:31-118): per-agent boxes in
    the agent's own frame + their uncertainties, for the box-alignment pose graph.  Returns three lists (one entry per
    agent): corners (K,8,3), boxes (K,7), uncertainty (K,ud) - or (None, None, None) when nothing passes the threshold."""
    n = cls_preds.shape[0]
    ud = unc_preds.shape[1] // cls_preds.shape[1]
    prob = torch.sigmoid(cls_preds.permute(0, 2, 3, 1).contiguous())
    unc = unc_preds.permute(0, 2, 3, 1).contiguous()
    batch_box3d = delta_to_boxes3d(reg_preds, anchor_box)
    mask = torch.gt(prob, params["target_args"]["score_threshold"])
    counts = [int(m.sum()) for m in mask]
    mask = mask.view(-1)
    boxes3d = batch_box3d.view(-1, 7)[mask].view(-1, 7)
    uncertainty = unc.view(-1, ud)[mask].view(-1, ud)
    scores = prob.view(-1)[mask]
    if dir_preds is not None and len(boxes3d) != 0:
        dir_offset = params["dir_args"]["dir_offset"]
        num_bins = params["dir_args"]["num_bins"]
        dcp = dir_preds.permute(0, 2, 3, 1).contiguous().reshape(-1, num_bins)[mask]
        dir_labels = torch.max(dcp, dim=-1)[1]
        period = 2 * np.pi / num_bins
        dir_rot = limit_period(boxes3d[..., 6] - dir_offset, 0, period)
        boxes3d[..., 6] = dir_rot + dir_offset + period * dir_labels.to(boxes3d.dtype)
        boxes3d[..., 6] = limit_period(boxes3d[..., 6], 0.5, 2 * np.pi)
    if len(boxes3d) == 0:
        return None, None, None
    corners = boxes_to_corners_3d(boxes3d, order=params["order"])         # no projection at stage 1 (:80-82)
    out_c, out_b, out_u, cur = [], [], [], 0
    for k in counts:
        c, b, s, u = corners[cur:cur + k], boxes3d[cur:cur + k], scores[cur:cur + k], uncertainty[cur:cur + k]
        keep = nms_rotated(c, s, params["nms_thresh"])
        out_c.append(c[keep]); out_b.append(b[keep]); out_u.append(u[keep])
        cur += k
    return out_c, out_b, out_u
