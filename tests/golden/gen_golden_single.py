"""Golden vectors of the UNMODIFIED reference single-agent `PointPillar` (BASELINE configs[0]:
opencood/models/point_pillar.py with BaseBEVBackbone, yaml opv2v/lidar_only_with_noise/pointpillar_single.yaml),
created through the reference's own yaml loader + model registry.  Build container only:

    python tests/golden/gen_golden_single.py     ->  tests/golden/model_single_plain.npz

Input: a batch of 3 independent single-agent frames (small canvas, see tests/golden_cases.py); voxel tensors from
oracle/voxelize.c (spconv absent: that stage is unpinned), everything downstream is the reference.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", ROOT]

import numpy as np
import torch

from coalign_b200 import synth
from tests.golden_cases import SMALL_RANGE, SMALL_VOXEL, single_case_inputs

from opencood.hypes_yaml import yaml_utils                      # noqa: E402  (reference)
from opencood.tools import train_utils                          # noqa: E402

REF_YAML = "/root/reference/opencood/hypes_yaml/opv2v/lidar_only_with_noise/pointpillar_single.yaml"
REF_YAML_UNC = "/root/reference/opencood/hypes_yaml/opv2v/lidar_only_with_noise/coalign/pointpillar_uncertainty.yaml"


def main(seed=5, n_frames=3):
    args = synth.single_args(SMALL_RANGE, SMALL_VOXEL)
    sd = synth.random_state_dict(args, seed, backbone="plain")
    hypes = yaml_utils.load_yaml(REF_YAML)
    margs = hypes["model"]["args"]
    margs["lidar_range"] = args["lidar_range"]
    margs["voxel_size"] = args["voxel_size"]
    margs["point_pillar_scatter"]["grid_size"] = args["point_pillar_scatter"]["grid_size"]
    assert hypes["model"]["core_method"] == "point_pillar"
    model = train_utils.create_model(hypes)                      # registry path (train_utils.py:113-146)
    ref_sd = model.state_dict()
    assert set(ref_sd.keys()) == set(sd.keys()), sorted(set(ref_sd) ^ set(sd))
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    model.eval()
    inp = single_case_inputs(n_frames, seed0=100 + seed)
    data = {"processed_lidar": {"voxel_features": torch.from_numpy(inp["voxel_features"]),
                                "voxel_coords": torch.from_numpy(inp["voxel_coords"]),
                                "voxel_num_points": torch.from_numpy(inp["voxel_num_points"])}}
    stages = {}
    hooks = [model.backbone.register_forward_hook(lambda m, i, o: stages.__setitem__("decoded", o["spatial_features_2d"])),
             model.shrink_conv.register_forward_hook(lambda m, i, o: stages.__setitem__("shrunk", o))]
    for li, blk in enumerate(model.backbone.blocks):
        hooks.append(blk.register_forward_hook(lambda m, i, o, li=li: stages.__setitem__(f"feat{li}", o)))
    with torch.no_grad():
        out = model(data)
    for h in hooks:
        h.remove()
    rec = {"seed": np.int64(seed), "n_frames": np.int64(n_frames), "voxel_coords": inp["voxel_coords"],
           "voxel_num_points": inp["voxel_num_points"],
           "feat0": stages["feat0"].numpy(), "feat1": stages["feat1"].numpy(),
           "feat2": stages["feat2"].numpy(), "decoded": stages["decoded"].numpy(),
           "shrunk": stages["shrunk"].numpy()}
    for k, v in out.items():
        rec[k] = v.numpy()
    np.savez(os.path.join(HERE, "model_single_plain.npz"), **rec)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in rec.items()})


def main_uncertainty(seed=7, n_frames=2):
    """Stage-1 detector with the uncertainty head (opencood/models/point_pillar_uncertainty.py, yaml
    coalign/pointpillar_uncertainty.yaml) -> tests/golden/model_single_uncertainty.npz"""
    args = synth.uncertainty_args(SMALL_RANGE, SMALL_VOXEL)
    sd = synth.random_state_dict(args, seed, backbone="plain")
    hypes = yaml_utils.load_yaml(REF_YAML_UNC)
    margs = hypes["model"]["args"]
    margs["lidar_range"] = args["lidar_range"]
    margs["voxel_size"] = args["voxel_size"]
    margs["point_pillar_scatter"]["grid_size"] = args["point_pillar_scatter"]["grid_size"]
    assert hypes["model"]["core_method"] == "point_pillar_uncertainty"
    assert margs["uncertainty_dim"] == args["uncertainty_dim"] and "shrink_header" not in margs
    model = train_utils.create_model(hypes)
    ref_sd = model.state_dict()
    assert set(ref_sd.keys()) == set(sd.keys()), sorted(set(ref_sd) ^ set(sd))
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    model.eval()
    inp = single_case_inputs(n_frames, seed0=100 + seed)
    data = {"processed_lidar": {"voxel_features": torch.from_numpy(inp["voxel_features"]),
                                "voxel_coords": torch.from_numpy(inp["voxel_coords"]),
                                "voxel_num_points": torch.from_numpy(inp["voxel_num_points"])}}
    stages = {}
    h = model.backbone.register_forward_hook(lambda m, i, o: stages.__setitem__("decoded", o["spatial_features_2d"]))
    with torch.no_grad():
        out = model(data)
    h.remove()
    rec = {"seed": np.int64(seed), "n_frames": np.int64(n_frames), "decoded": stages["decoded"].numpy()}
    for k, v in out.items():
        rec[k] = v.numpy()
    np.savez(os.path.join(HERE, "model_single_uncertainty.npz"), **rec)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in rec.items()})


if __name__ == "__main__":
    torch.set_num_threads(8)
    main()
    main_uncertainty()
