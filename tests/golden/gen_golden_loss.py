"""Golden vectors for the loss: loss terms and autograd gradients of the UNMODIFIED reference
`PointPillarLoss` (opencood/loss/point_pillar_loss.py) created through the reference's own loss registry.  Build container:

    python tests/golden/gen_golden_loss.py        ->  tests/golden/loss_*.npz
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", ROOT]

import numpy as np
import torch

# import-only stand-ins for modules the reference pulls in but never executes on this path (see gen_golden_post.py)
for name in ("open3d",):
    sys.modules[name] = types.ModuleType(name)
import matplotlib                                               # noqa: E402  (stub package in _stubs)
cm = types.ModuleType("matplotlib.cm")
cm.get_cmap = lambda name: types.SimpleNamespace(colors=np.zeros((256, 3)))
matplotlib.cm = cm
sys.modules["matplotlib.cm"] = cm
bo = types.ModuleType("opencood.utils.box_overlaps")
bo.bbox_overlaps = None
sys.modules["opencood.utils.box_overlaps"] = bo

from opencood.tools import train_utils                          # noqa: E402  (reference)
from coalign_b200 import synth                                  # noqa: E402

CASES = {
    "typical": dict(seed=3, n=2, H=12, W=20),
    "empty_sample": dict(seed=4, n=3, H=10, W=16, empty_samples=(1,)),      # pos_normalizer clamp(min=1)
    "float32_labels": dict(seed=5, n=1, H=8, W=12, dtype=np.float32),
}


def run_case(name, **kw):
    hypes = {"loss": {"core_method": "point_pillar_loss", "args": synth.loss_args()}}
    crit = train_utils.create_loss(hypes)                        # reference registry (train_utils.py:149-175)
    assert type(crit).__name__ == "PointPillarLoss"
    case = synth.loss_case(**kw)
    out = {k2: torch.from_numpy(case[k]).requires_grad_(True) for k, k2 in (("cls", "cls_preds"), ("reg", "reg_preds"), ("dir", "dir_preds"))}
    tgt = {"pos_equal_one": torch.from_numpy(case["pos"]), "neg_equal_one": torch.from_numpy(case["neg"]),
           "targets": torch.from_numpy(case["tgt"])}
    total = crit(out, tgt)
    total.backward()
    res = {"total_loss": np.float64(total.item())}
    for k in ("reg_loss", "cls_loss", "dir_loss"):
        res[k] = np.float64(crit.loss_dict[k])
    res["g_cls"] = out["cls_preds"].grad.numpy()
    res["g_reg"] = out["reg_preds"].grad.numpy()
    res["g_dir"] = out["dir_preds"].grad.numpy()
    np.savez_compressed(os.path.join(HERE, f"loss_{name}.npz"), **res)
    print(name, {k: float(res[k]) for k in ("total_loss", "reg_loss", "cls_loss", "dir_loss")}, total.dtype)


if __name__ == "__main__":
    for name, kw in CASES.items():
        run_case(name, **kw)
