"""Golden vectors for the TRAINING step: the UNMODIFIED reference model in `.train()` mode (batch-stat
BatchNorm) + the unmodified reference `PointPillarLoss`, created through the reference's yaml loader and registries, one
forward / backward on the small synthetic case.  Build container only:

    python tests/golden/gen_golden_train.py       ->  tests/golden/train_small.npz

Stored (a full gradient set is 52 MB, so per tensor): L2 norm of the gradient, its first 4 entries and 4 entries at
fixed strides; the loss terms; the head outputs; the running statistics three BatchNorm layers hold after the step.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", ROOT]

import numpy as np
import torch

sys.modules["open3d"] = types.ModuleType("open3d")
import matplotlib                                               # noqa: E402  (stub package in _stubs)
cm = types.ModuleType("matplotlib.cm")
cm.get_cmap = lambda name: types.SimpleNamespace(colors=np.zeros((256, 3)))
matplotlib.cm = cm
sys.modules["matplotlib.cm"] = cm
bo = types.ModuleType("opencood.utils.box_overlaps")
bo.bbox_overlaps = None
sys.modules["opencood.utils.box_overlaps"] = bo

from coalign_b200 import synth                                  # noqa: E402
from tests.golden_cases import SMALL_RANGE, small_case_inputs   # noqa: E402
from tests.golden.gen_golden import build_reference_model       # noqa: E402
from opencood.tools import train_utils                          # noqa: E402  (reference)

BN_PROBES = ("pillar_vfe.pfn_layers.0.norm", "backbone.resnet.layer0.0.bn1", "backbone.resnet.layer2.0.downsample.1",
             "backbone.deblocks.2.1")


def sample_index(numel):
    """4 fixed positions spread over a flattened tensor."""
    return np.array([(numel * k) // 5 for k in (1, 2, 3, 4)], dtype=np.int64)


def main(seed=3, record_len=(2, 3)):
    args = synth.make_args(SMALL_RANGE, [0.4, 0.4, 4])
    sd = synth.random_state_dict(args, seed)
    model = build_reference_model(args, sd, "att")
    model.train()                                                # train.py:109
    crit = train_utils.create_loss({"loss": {"core_method": "point_pillar_loss", "args": synth.loss_args()}})
    inp = small_case_inputs(list(record_len), seed0=100 + seed)
    data = {"processed_lidar": {"voxel_features": torch.from_numpy(inp["voxel_features"]),
                                "voxel_coords": torch.from_numpy(inp["voxel_coords"]),
                                "voxel_num_points": torch.from_numpy(inp["voxel_num_points"])},
            "record_len": torch.from_numpy(inp["record_len"]),
            "pairwise_t_matrix": torch.from_numpy(inp["pairwise_t_matrix"])}
    model.zero_grad()
    out = model(data)                                            # train.py:114
    B, _, H, W = out["cls_preds"].shape
    case = synth.loss_case(seed=seed, n=B, H=H, W=W, n_pos=6)
    label = {"pos_equal_one": torch.from_numpy(case["pos"]), "neg_equal_one": torch.from_numpy(case["neg"]),
             "targets": torch.from_numpy(case["tgt"])}
    loss = crit(out, label)                                      # train.py:116
    loss.backward()                                              # train.py:124
    res = {"seed": np.int64(seed), "record_len": np.asarray(record_len, np.int64),
           "total_loss": np.float64(loss.item())}
    for k in ("reg_loss", "cls_loss", "dir_loss"):
        res[k] = np.float64(crit.loss_dict[k])
    for k, v in out.items():
        res["out_" + k] = v.detach().numpy()
    n_param = 0
    for name, p in model.named_parameters():
        assert p.grad is not None, name
        g = p.grad.detach().double().flatten().numpy()
        res["gn_" + name] = np.float64(np.sqrt((g * g).sum()))
        res["g4_" + name] = g[:4].copy()
        res["gs_" + name] = g[sample_index(g.size)].copy()
        n_param += 1
    msd = model.state_dict()
    for pre in BN_PROBES:
        res["rm_" + pre] = msd[pre + ".running_mean"].numpy().copy()
        res["rv_" + pre] = msd[pre + ".running_var"].numpy().copy()
    np.savez_compressed(os.path.join(HERE, "train_small.npz"), **res)
    print("params with grad:", n_param, "losses:", {k: float(res[k]) for k in ("total_loss", "reg_loss", "cls_loss", "dir_loss")})


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
