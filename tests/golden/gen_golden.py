"""Generate golden vectors by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/gen_golden.py

Writes tests/golden/*.npz.  The fixtures pin ``oracle/coalign_oracle.py`` (tests/test_oracle_golden.py)
and are also compared directly with the CUDA path (tests/test_parity_gpu.py).

Import-only stubs for packages the reference imports but never executes on the CoAlign path
(icecream, matplotlib, shapely, pyquaternion, turtle) live in tests/golden/_stubs.
Voxelization is NOT part of the reference tree (spconv, absent) - voxel tensors fed to the reference
model come from oracle/voxelize.c ("parity unpinned" stage); everything downstream is the reference.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", ROOT]

import numpy as np
import torch

from coalign_b200 import synth
from tests.golden_cases import SMALL_RANGE, small_case_inputs

from opencood.hypes_yaml import yaml_utils                      # noqa: E402  (reference)
from opencood.tools import train_utils                          # noqa: E402
from opencood.utils.transformation_utils import normalize_pairwise_tfm, x_to_world, \
    get_pairwise_transformation                                  # noqa: E402
from opencood.models.sub_modules.torch_transformation_utils import warp_affine_simple   # noqa: E402
from opencood.models.fuse_modules.fusion_in_one import AttFusion, MaxFusion              # noqa: E402

REF_YAML = "/root/reference/opencood/hypes_yaml/opv2v/lidar_only_with_noise/coalign/pointpillar_coalign.yaml"

def build_reference_model(args, sd, fusion_method="att"):
    hypes = yaml_utils.load_yaml(REF_YAML)                       # unmodified yaml + parser
    margs = hypes["model"]["args"]
    margs["lidar_range"] = args["lidar_range"]
    margs["voxel_size"] = args["voxel_size"]
    margs["point_pillar_scatter"]["grid_size"] = args["point_pillar_scatter"]["grid_size"]
    margs["fusion_method"] = fusion_method
    model = train_utils.create_model(hypes)                      # registry path (train_utils.py:113-146)
    ref_sd = model.state_dict()
    assert set(ref_sd.keys()) == set(sd.keys()), (set(ref_sd) ^ set(sd))
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    model.load_state_dict(sd, strict=True)
    return model.eval()


def run_model_case(name, record_len, seed, fusion_method="att"):
    args = synth.make_args(SMALL_RANGE, [0.4, 0.4, 4])
    args["fusion_method"] = fusion_method
    sd = synth.random_state_dict(args, seed)
    model = build_reference_model(args, sd, fusion_method)
    inp = small_case_inputs(record_len, seed0=100 + seed)
    data = {"processed_lidar": {"voxel_features": torch.from_numpy(inp["voxel_features"]),
                                "voxel_coords": torch.from_numpy(inp["voxel_coords"]),
                                "voxel_num_points": torch.from_numpy(inp["voxel_num_points"])},
            "record_len": torch.from_numpy(inp["record_len"]),
            "pairwise_t_matrix": torch.from_numpy(inp["pairwise_t_matrix"])}
    stages = {}
    hooks = []

    def keep(key):
        def fn(_m, _i, o):
            stages.setdefault(key, []).append(o)
        return fn
    hooks.append(model.pillar_vfe.register_forward_hook(lambda m, i, o: stages.__setitem__("pillar_features", o["pillar_features"])))
    hooks.append(model.backbone.resnet.register_forward_hook(lambda m, i, o: stages.__setitem__("feats", o)))
    for i, f in enumerate(model.fusion_net):
        hooks.append(f.register_forward_hook(keep("fused")))
    if hasattr(model, "shrink_conv"):
        hooks.append(model.shrink_conv.register_forward_hook(lambda m, i, o: stages.__setitem__("shrunk", o)))
    with torch.no_grad():
        out = model(data)
    for h in hooks:
        h.remove()
    rec = {"record_len": inp["record_len"], "seed": np.int64(seed),
           "voxel_coords": inp["voxel_coords"], "voxel_num_points": inp["voxel_num_points"],
           "voxel_features_sum": np.float64(inp["voxel_features"].astype(np.float64).sum()),
           "pillar_features": stages["pillar_features"].numpy(),
           "shrunk": stages["shrunk"].numpy()}
    for i in range(3):
        rec[f"feat{i}"] = stages["feats"][i].numpy()
        rec[f"fused{i}"] = stages["fused"][i].numpy()
    for k, v in out.items():
        rec[k] = v.numpy()
    np.savez(os.path.join(HERE, name + ".npz"), **rec)
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in rec.items()})


def run_op_cases():
    """Small op-level vectors for A10/A11/A12 (+ max fusion), incl. identity and far-out-of-view warps."""
    g = torch.Generator().manual_seed(7)
    rec = {}
    # A10: normalize_pairwise_tfm on random rigid transforms, H,W like the model call (:108-109)
    rng = np.random.default_rng(7)
    poses = [np.array([rng.uniform(-30, 30), rng.uniform(-30, 30), 0, 0, rng.uniform(-180, 180), 0]) for _ in range(4)]
    base = {i: {"params": {"lidar_pose": list(p)}} for i, p in enumerate(poses)}
    pw = get_pairwise_transformation(base, 5, False)
    rec["poses"] = np.stack(poses)
    rec["pairwise"] = pw
    rec["affine_200_704"] = normalize_pairwise_tfm(torch.from_numpy(pw[None].copy()), 200, 704, 0.4).numpy()
    # A11: warp on odd sizes
    src = torch.randn(4, 6, 9, 13, generator=g)
    M = torch.tensor([[[1, 0, 0], [0, 1, 0]],
                      [[0.8, -0.5, 0.2], [0.6, 0.9, -0.3]],
                      [[-1.0, 0.05, 1.7], [0.02, -1.0, 0.4]],
                      [[1, 0, 5.0], [0, 1, 5.0]]], dtype=torch.float64)
    rec["warp_src"] = src.numpy()
    rec["warp_M"] = M.numpy()
    rec["warp_out"] = warp_affine_simple(src, M, (9, 13)).numpy()
    # A12: AttFusion / MaxFusion on (sumN=5: scenes of 3 and 2)
    x = torch.randn(5, 16, 10, 14, generator=g)
    rl = torch.tensor([3, 2])
    aff = torch.zeros(2, 5, 5, 2, 3, dtype=torch.float64)
    aff[..., 0, 0] = 1
    aff[..., 1, 1] = 1
    aff[0, 0, 1] = M[1]
    aff[0, 0, 2] = M[2]
    aff[1, 0, 1] = torch.tensor([[0.95, 0.3, -0.1], [-0.3, 0.95, 0.2]], dtype=torch.float64)
    rec["att_x"] = x.numpy()
    rec["att_affine"] = aff.numpy()
    rec["att_out"] = AttFusion(16)(x, rl, aff).numpy()
    rec["max_out"] = MaxFusion()(x, rl, aff).numpy()
    np.savez(os.path.join(HERE, "ops.npz"), **rec)
    print("ops", {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    torch.set_num_threads(8)
    run_op_cases()
    run_model_case("model_small_att", [3, 2], seed=1, fusion_method="att")
    run_model_case("model_small_single", [1], seed=2, fusion_method="att")
    run_model_case("model_small_max", [2, 3], seed=3, fusion_method="max")
