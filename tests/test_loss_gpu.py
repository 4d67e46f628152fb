"""SURVEY 8f row 2 (first piece): cb_pointpillar_loss through PointPillarLossB200 against the unmodified reference's loss
terms and autograd gradients (golden vectors from tests/golden/gen_golden_loss.py)."""
import os

import numpy as np
import pytest
import torch

from coalign_b200 import synth
from tests.test_loss_cpu import CASES, GOLD

pytestmark = pytest.mark.gpu


def run_cuda(case, requires_grad=False):
    from coalign_b200.loss import PointPillarLossB200
    crit = PointPillarLossB200(synth.loss_args())
    out = {"cls_preds": torch.from_numpy(case["cls"]).cuda().requires_grad_(requires_grad),
           "reg_preds": torch.from_numpy(case["reg"]).cuda().requires_grad_(requires_grad),
           "dir_preds": torch.from_numpy(case["dir"]).cuda().requires_grad_(requires_grad)}
    tgt = {"pos_equal_one": torch.from_numpy(case["pos"]).cuda(), "neg_equal_one": torch.from_numpy(case["neg"]).cuda(),
           "targets": torch.from_numpy(case["tgt"])}                      # labels may still be on the host
    total = crit(out, tgt)
    return crit, out, total


@pytest.mark.parametrize("name", list(CASES))
def test_loss_cuda_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, f"loss_{name}.npz"))
    crit, _, total = run_cuda(synth.loss_case(**CASES[name]))
    # float32 outputs of float64 sums; the reference's float32 branches differ by float32 rounding
    assert abs(float(total) - float(g["total_loss"])) <= 2e-6 * abs(float(g["total_loss"]))
    for k in ("total_loss", "reg_loss", "cls_loss", "dir_loss"):
        assert abs(crit.loss_dict[k] - float(g[k])) <= 2e-6 * abs(float(g[k])), (k, crit.loss_dict[k], float(g[k]))
    for k, gk in (("cls_preds", "g_cls"), ("reg_preds", "g_reg"), ("dir_preds", "g_dir")):
        np.testing.assert_allclose(crit.last_grads[k].cpu().numpy(), g[gk], rtol=2e-5, atol=2e-8, err_msg=k)


def test_loss_backward_through_autograd_and_full_size():
    """loss.backward() hands the kernel's gradients to whatever produced the head outputs; OPV2V-size maps
    (4 samples x 100 x 352 x 2 anchors) against the oracle; two runs are bit-identical (fixed-order reductions)."""
    from oracle import loss_oracle as LO
    case = synth.loss_case(seed=9, n=4, H=100, W=352, n_pos=40)
    crit, out, total = run_cuda(case, requires_grad=True)
    (total * 0.5).backward()
    losses, grads = LO.loss_and_grads(synth.loss_args(), case)
    assert abs(float(total.detach()) - losses["total_loss"]) <= 2e-6 * losses["total_loss"]
    for k, gk in (("cls_preds", "cls"), ("reg_preds", "reg"), ("dir_preds", "dir")):
        np.testing.assert_allclose(out[k].grad.cpu().numpy(), 0.5 * grads[gk], rtol=2e-5, atol=1e-9, err_msg=k)
    crit2, _, total2 = run_cuda(case)
    assert float(total2) == float(total.detach())
    for k in crit.last_grads:
        assert torch.equal(crit.last_grads[k], crit2.last_grads[k])
