"""GPU parity of the detection post-processing (cb_postprocess through the C ABI / the VoxelPostprocessorB200 mirror)
against the golden vectors of the unmodified reference and against the CPU oracle at full OPV2V size."""
import os

import numpy as np
import pytest
import torch

from tests import golden_cases as G
from tests.test_postprocess_cpu import CASES, GOLD, case_inputs

pytestmark = pytest.mark.gpu


def run_cuda(params, anchors, inp, with_dir=True):
    from coalign_b200.postprocess import VoxelPostprocessorB200
    pp = VoxelPostprocessorB200(params, train=False)
    data_dict = {"ego": {"transformation_matrix": torch.from_numpy(inp["tfm"]).cuda(),
                         "anchor_box": torch.from_numpy(anchors)}}
    out = {"ego": {"cls_preds": torch.from_numpy(inp["cls"]).cuda(), "reg_preds": torch.from_numpy(inp["reg"]).cuda()}}
    if with_dir:
        out["ego"]["dir_preds"] = torch.from_numpy(inp["dir"]).cuda()
    return pp.post_process(data_dict, out)


def check(boxes, scores, ref_boxes, ref_scores, what):
    assert boxes.shape == ref_boxes.shape, (what, boxes.shape, ref_boxes.shape)
    # float32 sigmoid / exp / sin / cos of the device differ from the host's by an ulp or two
    np.testing.assert_allclose(scores.cpu().numpy(), ref_scores, rtol=0, atol=3e-7, err_msg=what)
    np.testing.assert_allclose(boxes.cpu().numpy(), ref_boxes, rtol=0, atol=2e-5, err_msg=what)


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, f"post_{name}.npz"))
    params, anchors, inp = case_inputs(name)
    boxes, scores = run_cuda(params, anchors, inp)
    if not bool(g["has_result"]):
        assert boxes is None and scores is None
        return
    check(boxes, scores, g["boxes"], g["scores"], name)


@pytest.mark.parametrize("cls_bias,with_dir", [(-4.0, True), (-0.5, True), (-4.0, False)])
def test_full_size_vs_oracle(cls_bias, with_dir):
    """OPV2V head maps (100 x 352 x 2 anchors = 70400): a detection-like case and a >30k-candidate case (radix select)."""
    from oracle import postprocess_oracle as PO
    params = G.post_params()
    anchors = PO.generate_anchor_box(params)
    inp = G.post_case_inputs(params, anchors, seed=21, cls_bias=cls_bias, n_objects=60, yaw_deg=2.0, shift=(0.5, -0.25, 0.0))
    rb, rs = PO.post_process(params, torch.from_numpy(anchors), torch.from_numpy(inp["tfm"]), torch.from_numpy(inp["cls"]),
                             torch.from_numpy(inp["reg"]), torch.from_numpy(inp["dir"]) if with_dir else None)
    boxes, scores = run_cuda(params, anchors, inp, with_dir)
    check(boxes, scores, rb.numpy(), rs.numpy(), f"full size bias {cls_bias}")


def test_batched_call_equals_single_calls_and_properties():
    """n scenes per call == n single calls; kept boxes are sorted by score, mutually below the IoU threshold (checked with
    the oracle's polygon IoU) and inside gt_range."""
    from coalign_b200.postprocess import VoxelPostprocessorB200
    from oracle import postprocess_oracle as PO
    params = G.post_params(H_map=32, W_map=48)
    anchors = PO.generate_anchor_box(params)
    inp = G.post_case_inputs(params, anchors, seed=33, cls_bias=-2.0, n_scenes=3)
    pp = VoxelPostprocessorB200(params, train=False)
    cls, reg, dr = (torch.from_numpy(inp[k]).cuda() for k in ("cls", "reg", "dir"))
    tfm = torch.from_numpy(inp["tfm"]).cuda()
    batched = pp.post_process_batch(cls, reg, dr, anchors, tfm)
    for b in range(3):
        (sb, ss), = pp.post_process_batch(cls[b:b + 1], reg[b:b + 1], dr[b:b + 1], anchors, tfm)
        assert torch.equal(batched[b][0], sb) and torch.equal(batched[b][1], ss)
        s = ss.cpu().numpy()
        assert (np.diff(s) <= 0).all() and (s > params["target_args"]["score_threshold"]).all()
        c = sb.cpu().numpy()
        lr = np.asarray(params["gt_range"])
        assert ((c >= lr[:3]) & (c <= lr[3:])).all()
        for i in range(len(c) - 1):
            iou = PO.quad_iou_one_to_many(c[i, :4, :2], c[i + 1:, :4, :2])
            assert (iou <= params["nms_thresh"]).all()


@pytest.mark.parametrize("name", ["typical", "one_empty", "none"])
def test_stage1_cuda_matches_reference_golden(name):
    """cb_postprocess_stage1 through UncertaintyVoxelPostprocessorB200.post_process_stage1 against the
    unmodified reference's post_process_stage1 (corners, 7-parameter boxes and gathered uncertainties per agent)."""
    from coalign_b200.postprocess import UncertaintyVoxelPostprocessorB200
    from tests.test_postprocess_cpu import stage1_case
    g = np.load(os.path.join(GOLD, f"post_stage1_{name}.npz"))
    params, anchors, inp = stage1_case(name)
    pp = UncertaintyVoxelPostprocessorB200(params, train=False)
    out = {"cls_preds": torch.from_numpy(inp["cls"]).cuda(), "reg_preds": torch.from_numpy(inp["reg"]).cuda(),
           "unc_preds": torch.from_numpy(inp["unc"]).cuda(), "dir_preds": torch.from_numpy(inp["dir"]).cuda()}
    c, b, u = pp.post_process_stage1(out, torch.from_numpy(anchors))
    if not bool(g["has_result"]):
        assert c is None and b is None and u is None
        return
    assert len(c) == int(g["n_agents"])
    for a in range(len(c)):
        assert c[a].shape == g[f"corners{a}"].shape, (a, c[a].shape, g[f"corners{a}"].shape)
        np.testing.assert_allclose(c[a].cpu().numpy(), g[f"corners{a}"], rtol=0, atol=2e-5)
        np.testing.assert_allclose(b[a].cpu().numpy(), g[f"boxes{a}"], rtol=0, atol=2e-5)
        assert np.array_equal(u[a].cpu().numpy(), g[f"unc{a}"])
