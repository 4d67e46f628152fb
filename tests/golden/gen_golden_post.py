"""Golden vectors for the detection post-processing, produced by the UNMODIFIED reference
`VoxelPostprocessor` imported from /root/reference.  Build container only:

    python tests/golden/gen_golden_post.py        ->  tests/golden/post_*.npz

Third-party modules the reference imports but that are absent here are replaced before the import:
  * open3d, matplotlib(.cm/.pyplot), icecream, pyquaternion, turtle: import-only stubs (never executed on this path);
  * opencood.utils.box_overlaps (Cython, only used by generate_label): import-only stub;
  * shapely.geometry.Polygon: a FUNCTIONAL stand-in backed by oracle/rotated_iou.c (area / intersection / union of
    convex polygons).  The reference's nms_rotated -> common_utils.compute_iou runs unmodified on top of it, so the
    fixtures pin everything except the polygon IoU itself (parity unpinned there: shapely/GEOS absent).
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", ROOT]

import numpy as np
import torch

from oracle import postprocess_oracle as PO


class _Area:
    def __init__(self, area):
        self.area = area


class Polygon:
    """Convex-polygon stand-in for shapely.geometry.Polygon (only what common_utils.compute_iou touches)."""

    def __init__(self, pts):
        self.pts = np.asarray([(float(x), float(y)) for x, y in pts], dtype=np.float64)

    @property
    def area(self):
        x, y = self.pts[:, 0], self.pts[:, 1]
        return 0.5 * abs(float(np.dot(x, np.roll(y, -1)) - np.dot(np.roll(x, -1), y)))

    def intersection(self, other):
        return _Area(PO.convex_intersection_area(self.pts, other.pts))

    def union(self, other):
        return _Area(self.area + other.area - PO.convex_intersection_area(self.pts, other.pts))


geom = types.ModuleType("shapely.geometry")
geom.Polygon = Polygon
geom.Point = geom.MultiPoint = object
shp = types.ModuleType("shapely")
shp.geometry = geom
sys.modules["shapely"] = shp
sys.modules["shapely.geometry"] = geom
o3d = types.ModuleType("open3d")
sys.modules["open3d"] = o3d
import matplotlib                                               # noqa: E402  (stub package in _stubs)
cm = types.ModuleType("matplotlib.cm")
cm.get_cmap = lambda name: types.SimpleNamespace(colors=np.zeros((256, 3)))
matplotlib.cm = cm
sys.modules["matplotlib.cm"] = cm
bo = types.ModuleType("opencood.utils.box_overlaps")
bo.bbox_overlaps = None
sys.modules["opencood.utils.box_overlaps"] = bo

from opencood.data_utils.post_processor.voxel_postprocessor import VoxelPostprocessor   # noqa: E402  (reference)

from opencood.data_utils.post_processor.uncertainty_voxel_postprocessor import UncertaintyVoxelPostprocessor   # noqa: E402

from tests.golden_cases import post_params, post_case_inputs, stage1_case_inputs   # noqa: E402


def run_case(name, H, W, **kw):
    params = post_params(H_map=H, W_map=W)
    pp = VoxelPostprocessor(params, train=False)
    anchors = pp.generate_anchor_box()
    assert anchors.shape == (H, W, 2, 7), anchors.shape
    inp = post_case_inputs(params, anchors, **kw)
    data_dict = {"ego": {"transformation_matrix": torch.from_numpy(inp["tfm"]),
                         "anchor_box": torch.from_numpy(anchors)}}
    out_dict = {"ego": {"cls_preds": torch.from_numpy(inp["cls"]), "reg_preds": torch.from_numpy(inp["reg"]),
                        "dir_preds": torch.from_numpy(inp["dir"])}}
    boxes, scores = pp.post_process(data_dict, out_dict)
    res = {"anchors": anchors, "has_result": np.array(boxes is not None)}
    if boxes is not None:
        res["boxes"] = boxes.numpy()
        res["scores"] = scores.numpy()
    # intermediate stages through the reference's own helper functions
    res["decoded"] = VoxelPostprocessor.delta_to_boxes3d(torch.from_numpy(inp["reg"]), torch.from_numpy(anchors)).numpy()
    np.savez_compressed(os.path.join(HERE, f"post_{name}.npz"), **res)
    n = 0 if boxes is None else boxes.shape[0]
    thr = params["target_args"]["score_threshold"]
    n_cand = int((1 / (1 + np.exp(-inp["cls"])) > thr).sum())
    print(f"{name}: candidates {n_cand} -> kept {n}")


def run_stage1_case(name, H, W, n_agents, **kw):
    """UncertaintyVoxelPostprocessor.post_process_stage1 of the UNMODIFIED reference -> post_stage1_<name>.npz"""
    params = post_params(H_map=H, W_map=W)
    pp = UncertaintyVoxelPostprocessor(params, train=False)
    anchors = pp.generate_anchor_box()
    inp = stage1_case_inputs(params, anchors, n_agents=n_agents, **kw)
    out = {"cls_preds": torch.from_numpy(inp["cls"]), "reg_preds": torch.from_numpy(inp["reg"]),
           "unc_preds": torch.from_numpy(inp["unc"]), "dir_preds": torch.from_numpy(inp["dir"])}
    corners, boxes, unc = pp.post_process_stage1(out, torch.from_numpy(anchors))
    res = {"has_result": np.array(corners is not None), "n_agents": np.int64(n_agents)}
    if corners is not None:
        for b in range(n_agents):
            res[f"corners{b}"] = corners[b].numpy()
            res[f"boxes{b}"] = boxes[b].numpy()
            res[f"unc{b}"] = unc[b].numpy()
    np.savez_compressed(os.path.join(HERE, f"post_stage1_{name}.npz"), **res)
    print(f"stage1 {name}: kept", None if corners is None else [int(c.shape[0]) for c in corners])


if __name__ == "__main__":
    PO.build_c()
    run_stage1_case("typical", seed=11, H=24, W=40, n_agents=3, cls_bias=-3.0)
    run_stage1_case("one_empty", seed=12, H=16, W=24, n_agents=3, cls_bias=-4.0, n_objects=3, empty_agents=(1,))
    run_stage1_case("none", seed=13, H=16, W=24, n_agents=2, cls_bias=-12.0, n_objects=0)        # -> (None, None, None)
    run_case("typical", seed=1, H=24, W=40, cls_bias=-3.0)                 # a few dozen candidates, clustered
    run_case("many", seed=2, H=32, W=48, cls_bias=-0.3)                    # > 1000 candidates: top-1000 truncation
    run_case("none", seed=3, H=16, W=24, cls_bias=-12.0, n_objects=0)      # nothing above the threshold -> (None, None)
    run_case("filtered", seed=5, H=16, W=24, cls_bias=-3.0, shift=(0.0, 0.0, 5.0))   # candidates, all fail the z filter
    run_case("posed", seed=4, H=24, W=40, cls_bias=-2.5, yaw_deg=30.0, shift=(3.0, -2.0, 0.1))   # non-identity T
