"""Inputs of the golden cases (shared by tests/golden/gen_golden.py and the tests).

Deterministic from seeds: synthetic scenes (coalign_b200.synth) -> oracle voxelizer -> reference-format
batch dict.  No reference import here, so it runs on the GPU box too.
"""
import numpy as np

from coalign_b200 import synth
from oracle import voxelize_np as V

SMALL_RANGE = [-11.2, -4.8, -3, 11.2, 4.8, 1]      # -> nx=56, ny=24 (levels 28x12, 14x6, 7x3: odd sizes covered)
SMALL_VOXEL = [0.4, 0.4, 4]


def small_args(fusion_method="att"):
    a = synth.make_args(SMALL_RANGE, SMALL_VOXEL)
    a["fusion_method"] = fusion_method
    return a


def scenes_to_batch(scenes, lidar_range, voxel_size, max_pts=32, max_voxels=70000):
    per_agent, pws = [], []
    for sc in scenes:
        for pts in sc["points"]:
            per_agent.append(V.voxelize_c(pts, lidar_range, voxel_size, max_pts, max_voxels))
        pws.append(sc["pairwise_t_matrix"])
    vf, vc, vn = V.collate(per_agent)
    return {"voxel_features": vf, "voxel_coords": vc, "voxel_num_points": vn,
            "record_len": np.asarray([len(sc["points"]) for sc in scenes], np.int64),
            "pairwise_t_matrix": np.stack(pws)}


def small_case_scenes(record_len, seed0, n_points=1800):
    return [synth.make_scene(seed0 + b, n, n_points, SMALL_RANGE, max_cav=5, pose_noise=True,
                             spread=5.0, sigma=5.0) for b, n in enumerate(record_len)]


def small_case_inputs(record_len, seed0, n_points=1800):
    return scenes_to_batch(small_case_scenes(record_len, seed0, n_points), SMALL_RANGE, SMALL_VOXEL)


def to_torch_batch(inp):
    import torch
    return {"processed_lidar": {"voxel_features": torch.from_numpy(inp["voxel_features"]),
                                "voxel_coords": torch.from_numpy(inp["voxel_coords"]),
                                "voxel_num_points": torch.from_numpy(inp["voxel_num_points"])},
            "record_len": torch.from_numpy(inp["record_len"]),
            "pairwise_t_matrix": torch.from_numpy(inp["pairwise_t_matrix"])}


# ------------------------------------------------------------------------------------------
# detection post-processing cases
# ------------------------------------------------------------------------------------------
post_params = synth.post_params


def post_case_inputs(params, anchors, seed, cls_bias=-3.0, n_objects=12, yaw_deg=0.0, shift=(0.0, 0.0, 0.0),
                     n_scenes=1):
    """Synthetic head outputs with detection-like structure: background logits N(cls_bias, 1.5), plus `n_objects` blobs of
    high-score anchors (overlapping boxes for the NMS to suppress); regression deltas N(0, 0.3); direction logits N(0,1).
    Returns float32 arrays cls (n,2,H,W), reg (n,14,H,W), dir (n,4,H,W) and the 4x4 cav->ego matrix tfm."""
    H, W, A = anchors.shape[:3]
    rng = np.random.default_rng(seed)
    cls = rng.normal(cls_bias, 1.5, (n_scenes, A, H, W))
    for b in range(n_scenes):
        for _ in range(n_objects):
            h0, w0 = int(rng.integers(1, H - 1)), int(rng.integers(1, W - 1))
            cls[b, :, h0 - 1:h0 + 2, w0 - 1:w0 + 2] += rng.uniform(2.0, 6.0, (A, 3, 3))
    reg = rng.normal(0.0, 0.3, (n_scenes, 7 * A, H, W))
    dr = rng.normal(0.0, 1.0, (n_scenes, 2 * A, H, W))
    c, s = np.cos(np.deg2rad(yaw_deg)), np.sin(np.deg2rad(yaw_deg))
    tfm = np.eye(4)
    tfm[:2, :2] = [[c, -s], [s, c]]
    tfm[:3, 3] = shift
    return {"cls": cls.astype(np.float32), "reg": reg.astype(np.float32), "dir": dr.astype(np.float32),
            "tfm": tfm.astype(np.float32)}


def stage1_case_inputs(params, anchors, seed, n_agents=3, uncertainty_dim=3, empty_agents=(), **kw):
    """Head outputs of the stage-1 uncertainty detector for `n_agents` agents: post_case_inputs + unc_preds
    (n, uncertainty_dim*A, H, W) ~ N(-1, 0.5) (log-variances); `empty_agents` get no anchor above the threshold."""
    inp = post_case_inputs(params, anchors, seed, n_scenes=n_agents, **kw)
    for b in empty_agents:
        inp["cls"][b] = -20.0
    H, W, A = anchors.shape[:3]
    rng = np.random.default_rng(seed + 7000)
    inp["unc"] = rng.normal(-1.0, 0.5, (n_agents, uncertainty_dim * A, H, W)).astype(np.float32)
    return inp


def single_case_inputs(n_frames, seed0, n_points=1800):
    """A batch of `n_frames` independent single-agent frames (the single-agent `point_pillar` model): reference collate
    format with the batch index in voxel_coords[:, 0]."""
    scenes = [synth.make_scene(seed0 + b, 1, n_points, SMALL_RANGE, max_cav=5, spread=5.0, sigma=5.0) for b in range(n_frames)]
    b = scenes_to_batch(scenes, SMALL_RANGE, SMALL_VOXEL)
    b["points"] = [sc["points"][0] for sc in scenes]
    return b
