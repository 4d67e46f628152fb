"""Host-side mirror of the reference's ``VoxelPostprocessor`` for the CoAlign path:
same constructor arguments (the yaml ``postprocess`` block after ``load_point_pillar_params``), same
``generate_anchor_box()`` and ``post_process(data_dict, output_dict)`` signatures, return values and ``(None, None)``
behaviour as /root/reference/opencood/data_utils/post_processor/voxel_postprocessor.py:25-82,243-402 - but decoding,
filters, top-k and the rotated NMS run in ``libcoalign_b200.so`` (``cb_postprocess``) instead of torch ops + a Python
loop over shapely polygons.  No CPU fallback: a missing library / non-B200 device raises.

Scope: intermediate fusion (``data_dict`` holds only ``'ego'``, as CoAlign's dataset produces it -
opencood/data_utils/datasets/intermediate_fusion_dataset.py); late fusion (several cavs per call) is not on this path.
``post_process_batch`` is the batched entry point (n scenes per call, one launch sequence).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib

TOP_K = 1000          # box_utils.nms_rotated: `top = 1000`


class VoxelPostprocessorB200:
    def __init__(self, anchor_params: dict, train: bool = False):
        self.params = anchor_params
        self.train = train
        self.bbx_dict = {}
        self.anchor_num = self.params["anchor_args"]["num"]
        self._ws: Optional[torch.Tensor] = None
        self._anchor_cache: Dict[tuple, torch.Tensor] = {}

    # voxel_postprocessor.py:30-82 - host-side numpy, executed once per run by the dataset
    def generate_anchor_box(self) -> np.ndarray:
        aa = self.params["anchor_args"]
        W, H = aa["W"], aa["H"]
        l, w, h = aa["l"], aa["w"], aa["h"]
        r = aa["r"]
        assert self.anchor_num == len(r)
        r = [math.radians(ele) for ele in r]
        vh, vw = aa["vh"], aa["vw"]
        xrange = [aa["cav_lidar_range"][0], aa["cav_lidar_range"][3]]
        yrange = [aa["cav_lidar_range"][1], aa["cav_lidar_range"][4]]
        feature_stride = aa["feature_stride"] if "feature_stride" in aa else 2
        x = np.linspace(xrange[0] + vw, xrange[1] - vw, W // feature_stride)
        y = np.linspace(yrange[0] + vh, yrange[1] - vh, H // feature_stride)
        cx, cy = np.meshgrid(x, y)
        cx = np.tile(cx[..., np.newaxis], self.anchor_num)
        cy = np.tile(cy[..., np.newaxis], self.anchor_num)
        cz = np.ones_like(cx) * -1.0
        w = np.ones_like(cx) * w
        l = np.ones_like(cx) * l
        h = np.ones_like(cx) * h
        r_ = np.ones_like(cx)
        for i in range(self.anchor_num):
            r_[..., i] = r[i]
        if self.params["order"] == "hwl":
            return np.stack([cx, cy, cz, h, w, l, r_], axis=-1)
        if self.params["order"] == "lhw":
            return np.stack([cx, cy, cz, l, h, w, r_], axis=-1)
        raise SystemExit("Unknown bbx order.")

    # ------------------------------------------------------------------ CUDA path
    def _device_anchors(self, anchor_box, device) -> torch.Tensor:
        """Device copy of the anchors, cached by CONTENT (a strided 257-element sample + shape + the total): an address
        can be reused by a different array, and per-batch collated anchor tensors are new objects with the same values."""
        if isinstance(anchor_box, np.ndarray):
            anchor_box = torch.from_numpy(anchor_box)
        flat = anchor_box.reshape(-1)
        step = max(1, flat.numel() // 257)
        sample = flat[::step].detach().to("cpu", torch.float64)
        key = (tuple(anchor_box.shape), str(device), float(flat.detach().double().sum()), tuple(sample.tolist()))
        t = self._anchor_cache.get(key)
        if t is None:
            t = anchor_box.to(device=device, dtype=torch.float32).contiguous()     # `.float()` of delta_to_boxes3d
            self._anchor_cache = {key: t}
        return t

    @torch.no_grad()
    def post_process_batch(self, cls_preds: torch.Tensor, reg_preds: torch.Tensor, dir_preds: Optional[torch.Tensor],
                           anchor_box, transformation_matrix: torch.Tensor, sync: bool = True):
        """n scenes per call.  Returns a list of (pred_box3d_tensor (K,8,3), scores (K,)) / (None, None) per scene, or -
        with sync=False - the raw device buffers (boxes (n,TOP_K,8,3), scores (n,TOP_K), counts (n,2)) without any
        host synchronisation."""
        lib = _lib.load(check_device=True)
        dev = cls_preds.device
        if dev.type != "cuda":
            raise RuntimeError("coalign_b200 post-processing needs CUDA tensors (no CPU fallback)")
        n, A, H, W = cls_preds.shape
        if A != self.anchor_num or reg_preds.shape != (n, 7 * A, H, W):
            raise ValueError("cls_preds / reg_preds shapes do not match the anchor configuration")
        num_bins = 0
        if dir_preds is not None:
            num_bins = int(self.params["dir_args"]["num_bins"])
            if dir_preds.shape != (n, num_bins * A, H, W):
                raise ValueError("dir_preds shape does not match dir_args.num_bins")
        anchors = self._device_anchors(anchor_box, dev)
        if tuple(anchors.shape) != (H, W, A, 7):
            raise ValueError("anchor_box must be (H, W, anchor_num, 7)")
        cls_c, reg_c = cls_preds.float().contiguous(), reg_preds.float().contiguous()
        dir_c = dir_preds.float().contiguous() if dir_preds is not None else None
        tfm = transformation_matrix.to(device=dev, dtype=torch.float32).reshape(-1, 4, 4)
        if tfm.shape[0] == 1 and n > 1:
            tfm = tfm.expand(n, 4, 4)
        tfm = tfm.contiguous()
        if tfm.shape[0] != n:
            raise ValueError("transformation_matrix must be (4,4) or (n,4,4)")
        need = int(lib.cb_postprocess_workspace_bytes(n, H, W, A))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        boxes = torch.empty(n, TOP_K, 8, 3, dtype=torch.float32, device=dev)
        scores = torch.empty(n, TOP_K, dtype=torch.float32, device=dev)
        counts = torch.zeros(n, 2, dtype=torch.int32, device=dev)
        gt_range = np.asarray(self.params["gt_range"], dtype=np.float64)
        dir_offset = float(self.params["dir_args"]["dir_offset"]) if dir_preds is not None else 0.0
        _lib.check(lib.cb_postprocess(cls_c.data_ptr(), reg_c.data_ptr(), dir_c.data_ptr() if dir_c is not None else None,
                                      n, H, W, A, num_bins, anchors.data_ptr(), tfm.data_ptr(),
                                      float(self.params["target_args"]["score_threshold"]), dir_offset,
                                      float(self.params["nms_thresh"]), gt_range.ctypes.data,
                                      1 if self.params["order"] == "hwl" else 0, TOP_K,
                                      boxes.data_ptr(), scores.data_ptr(), counts.data_ptr(),
                                      self._ws.data_ptr(), self._ws.numel(),
                                      torch.cuda.current_stream(dev).cuda_stream), "cb_postprocess")
        if not sync:
            return boxes, scores, counts
        cnt = counts.cpu().numpy()                    # the one device->host sync: result sizes
        out: List[Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]] = []
        for b in range(n):
            if cnt[b, 1] == 0:                        # nothing above the score threshold (voxel_postprocessor.py:366-367)
                out.append((None, None))
            else:
                k = int(cnt[b, 0])
                out.append((boxes[b, :k], scores[b, :k]))
        return out

    def post_process(self, data_dict: dict, output_dict: dict):
        """Same contract as VoxelPostprocessor.post_process (voxel_postprocessor.py:243-402)."""
        if len(data_dict) != 1:
            raise NotImplementedError("late fusion (several cavs per call) is not on the CoAlign path")
        (cav_id, cav_content), = data_dict.items()
        assert cav_id in output_dict
        od = output_dict[cav_id]
        cls = od["psm"] if "psm" in od else od["cls_preds"]
        reg = od["rm"] if "rm" in output_dict else od["reg_preds"]          # sic: the reference tests output_dict here
        dm = od["dm"] if "dm" in output_dict else od.get("dir_preds")
        if "iou_preds" in od:
            raise NotImplementedError("iou_preds score rectification is not part of the CoAlign heads")
        assert cls.shape[0] == 1                                            # voxel_postprocessor.py:317
        (boxes, scores), = self.post_process_batch(cls, reg, dm, cav_content["anchor_box"],
                                                   cav_content["transformation_matrix"])
        return boxes, scores


class UncertaintyVoxelPostprocessorB200(VoxelPostprocessorB200):
    """Mirror of the reference's ``UncertaintyVoxelPostprocessor``
    (/root/reference/opencood/data_utils/post_processor/uncertainty_voxel_postprocessor.py:27-118) for the stage-1 detector
    (``point_pillar_uncertainty``): ``post_process_stage1`` turns the per-agent head outputs into the boxes +
    uncertainties that ``pose_graph_pre_calc.py`` stores for the box-alignment pose graph.  Decode, direction fix and the
    rotated NMS of every agent run in one ``cb_postprocess_stage1`` launch sequence; the uncertainties of the kept
    anchors are gathered on the device from the returned anchor indices."""

    @torch.no_grad()
    def post_process_stage1(self, stage1_output_dict: dict, anchor_box):
        lib = _lib.load(check_device=True)
        cls_preds, reg_preds = stage1_output_dict["cls_preds"], stage1_output_dict["reg_preds"]
        unc_preds = stage1_output_dict["unc_preds"]
        dir_preds = stage1_output_dict.get("dir_preds")
        dev = cls_preds.device
        if dev.type != "cuda":
            raise RuntimeError("coalign_b200 post-processing needs CUDA tensors (no CPU fallback)")
        n, A, H, W = cls_preds.shape
        if A != self.anchor_num or reg_preds.shape != (n, 7 * A, H, W) or unc_preds.shape[1] % A:
            raise ValueError("cls_preds / reg_preds / unc_preds shapes do not match the anchor configuration")
        ud = unc_preds.shape[1] // A                                         # :42
        num_bins = 0
        if dir_preds is not None:
            num_bins = int(self.params["dir_args"]["num_bins"])
            if dir_preds.shape != (n, num_bins * A, H, W):
                raise ValueError("dir_preds shape does not match dir_args.num_bins")
        anchors = self._device_anchors(anchor_box, dev)
        if tuple(anchors.shape) != (H, W, A, 7):
            raise ValueError("anchor_box must be (H, W, anchor_num, 7)")
        cls_c, reg_c = cls_preds.float().contiguous(), reg_preds.float().contiguous()
        dir_c = dir_preds.float().contiguous() if dir_preds is not None else None
        need = int(lib.cb_postprocess_workspace_bytes(n, H, W, A))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        corners = torch.empty(n, TOP_K, 8, 3, dtype=torch.float32, device=dev)
        boxes7 = torch.empty(n, TOP_K, 7, dtype=torch.float32, device=dev)
        index = torch.zeros(n, TOP_K, dtype=torch.int32, device=dev)
        scores = torch.empty(n, TOP_K, dtype=torch.float32, device=dev)
        counts = torch.zeros(n, 2, dtype=torch.int32, device=dev)
        dir_offset = float(self.params["dir_args"]["dir_offset"]) if dir_preds is not None else 0.0
        _lib.check(lib.cb_postprocess_stage1(cls_c.data_ptr(), reg_c.data_ptr(),
                                             dir_c.data_ptr() if dir_c is not None else None, n, H, W, A, num_bins,
                                             anchors.data_ptr(), float(self.params["target_args"]["score_threshold"]),
                                             dir_offset, float(self.params["nms_thresh"]),
                                             1 if self.params["order"] == "hwl" else 0, TOP_K,
                                             corners.data_ptr(), boxes7.data_ptr(), index.data_ptr(), scores.data_ptr(),
                                             counts.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                             torch.cuda.current_stream(dev).cuda_stream), "cb_postprocess_stage1")
        cnt = counts.cpu().numpy()                                           # the one device->host sync: result sizes
        if int(cnt[:, 1].sum()) == 0:                                        # no anchor of any agent above the threshold (:87-88)
            return None, None, None
        # uncertainty of anchor idx = (h*W + w)*A + a, component k: unc_preds[n, a*ud + k, h, w]   (:45,55)
        unc_flat = unc_preds.float().permute(0, 2, 3, 1).reshape(n, H * W * A, ud)
        out_c, out_b, out_u = [], [], []
        for b in range(n):
            k = int(cnt[b, 0])
            out_c.append(corners[b, :k])
            out_b.append(boxes7[b, :k])
            out_u.append(unc_flat[b][index[b, :k].long()])
        return out_c, out_b, out_u
