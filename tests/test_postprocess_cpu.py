"""CPU tests of the detection post-processing oracle: oracle/postprocess_oracle.py against the golden
vectors of the unmodified reference (tests/golden/gen_golden_post.py), analytic known-answer cases for the polygon IoU
restatement (oracle/rotated_iou.c; shapely absent -> parity unpinned there), and the host-side mirror's anchors."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import postprocess_oracle as PO
from tests import golden_cases as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {"typical": dict(seed=1, H=24, W=40, cls_bias=-3.0), "many": dict(seed=2, H=32, W=48, cls_bias=-0.3),
         "none": dict(seed=3, H=16, W=24, cls_bias=-12.0, n_objects=0),
         "filtered": dict(seed=5, H=16, W=24, cls_bias=-3.0, shift=(0.0, 0.0, 5.0)),
         "posed": dict(seed=4, H=24, W=40, cls_bias=-2.5, yaw_deg=30.0, shift=(3.0, -2.0, 0.1))}


def case_inputs(name):
    kw = dict(CASES[name])
    H, W = kw.pop("H"), kw.pop("W")
    params = G.post_params(H_map=H, W_map=W)
    anchors = PO.generate_anchor_box(params)
    return params, anchors, G.post_case_inputs(params, anchors, **kw)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, f"post_{name}.npz"))
    params, anchors, inp = case_inputs(name)
    assert np.array_equal(anchors, g["anchors"])                       # float64, same numpy expressions
    dec = PO.delta_to_boxes3d(torch.from_numpy(inp["reg"]), torch.from_numpy(anchors)).numpy()
    assert np.array_equal(dec, g["decoded"])
    boxes, scores = PO.post_process(params, torch.from_numpy(anchors), torch.from_numpy(inp["tfm"]),
                                    torch.from_numpy(inp["cls"]), torch.from_numpy(inp["reg"]), torch.from_numpy(inp["dir"]))
    if not bool(g["has_result"]):
        assert boxes is None and scores is None
        return
    assert boxes.shape == g["boxes"].shape, (boxes.shape, g["boxes"].shape)
    assert np.array_equal(scores.numpy(), g["scores"])                 # same picks in the same order
    np.testing.assert_allclose(boxes.numpy(), g["boxes"], rtol=0, atol=1e-6)


def test_polygon_iou_known_answers():
    sq = np.array([[0, 0], [2, 0], [2, 2], [0, 2]], float)
    assert PO.convex_intersection_area(sq, sq) == pytest.approx(4.0, abs=1e-12)
    assert PO.convex_intersection_area(sq, sq + [1, 0]) == pytest.approx(2.0, abs=1e-12)
    assert PO.convex_intersection_area(sq, sq + [2, 0]) == pytest.approx(0.0, abs=1e-12)      # touching edge
    assert PO.convex_intersection_area(sq, sq + [5, 5]) == 0.0
    assert PO.convex_intersection_area(sq, sq[::-1] + [1, 1]) == pytest.approx(1.0, abs=1e-12)  # clockwise input
    # unit square rotated by 45 degrees about the centre of a 2x2 square: fully inside -> its own area
    c, s = math.cos(math.pi / 4), math.sin(math.pi / 4)
    u = np.array([[-.5, -.5], [.5, -.5], [.5, .5], [-.5, .5]]) @ np.array([[c, s], [-s, c]]) + [1, 1]
    assert PO.convex_intersection_area(sq, u) == pytest.approx(1.0, abs=1e-12)
    # two congruent squares, one rotated 45 degrees about the common centre: regular octagon, area 8(sqrt2-1) r^2 ... r=1
    big = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], float)
    rot = big @ np.array([[c, s], [-s, c]])
    assert PO.convex_intersection_area(big, rot) == pytest.approx(8 * (math.sqrt(2) - 1), abs=1e-12)
    iou = PO.quad_iou_one_to_many(sq, np.stack([sq, sq + [1, 0], sq + [5, 5]]))
    np.testing.assert_allclose(iou, [1.0, 2.0 / 6.0, 0.0], atol=1e-7)
    assert iou.dtype == np.float32


def test_host_mirror_anchors_and_abi_symbols():
    from coalign_b200 import _lib
    from coalign_b200.postprocess import VoxelPostprocessorB200
    params = G.post_params()                                           # OPV2V: 100 x 352 x 2 anchors
    pp = VoxelPostprocessorB200(params, train=False)
    a = pp.generate_anchor_box()
    assert a.shape == (100, 352, 2, 7) and a.dtype == np.float64
    assert np.array_equal(a, PO.generate_anchor_box(params))
    lib = _lib.load()                                                  # symbols only; no compute without a GPU
    assert hasattr(lib, "cb_postprocess") and hasattr(lib, "cb_postprocess_workspace_bytes")


STAGE1_CASES = {
    "typical": dict(seed=11, H=24, W=40, n_agents=3, cls_bias=-3.0),
    "one_empty": dict(seed=12, H=16, W=24, n_agents=3, cls_bias=-4.0, n_objects=3, empty_agents=(1,)),
    "none": dict(seed=13, H=16, W=24, n_agents=2, cls_bias=-12.0, n_objects=0),
}


def stage1_case(name):
    kw = dict(STAGE1_CASES[name])
    H, W = kw.pop("H"), kw.pop("W")
    params = G.post_params(H_map=H, W_map=W)
    anchors = PO.generate_anchor_box(params)
    return params, anchors, G.stage1_case_inputs(params, anchors, **kw)


@pytest.mark.parametrize("name", list(STAGE1_CASES))
def test_stage1_oracle_matches_reference_golden(name):
    """UncertaintyVoxelPostprocessor.post_process_stage1 (boxes + uncertainties for the pose graph) -
    oracle restatement against the unmodified reference (tests/golden/gen_golden_post.py::run_stage1_case)."""
    g = np.load(os.path.join(GOLD, f"post_stage1_{name}.npz"))
    params, anchors, inp = stage1_case(name)
    c, b, u = PO.post_process_stage1(params, torch.from_numpy(anchors), torch.from_numpy(inp["cls"]),
                                     torch.from_numpy(inp["reg"]), torch.from_numpy(inp["unc"]), torch.from_numpy(inp["dir"]))
    if not bool(g["has_result"]):
        assert c is None and b is None and u is None
        return
    assert len(c) == int(g["n_agents"])
    for a in range(len(c)):
        assert c[a].shape == g[f"corners{a}"].shape, (a, c[a].shape)
        np.testing.assert_allclose(c[a].numpy(), g[f"corners{a}"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(b[a].numpy(), g[f"boxes{a}"], rtol=0, atol=1e-6)
        assert np.array_equal(u[a].numpy(), g[f"unc{a}"])              # a pure gather: same picks in the same order
    if name == "one_empty":
        assert c[1].shape[0] == 0 and c[0].shape[0] > 0
