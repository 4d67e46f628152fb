"""Synthetic OPV2V/DAIR-shape inputs and random-init weights (no dataset / checkpoint is reachable).

Follows SURVEY.md 8(d).  numpy/torch-CPU only; used by bench.py, smoke(), the tests and the
golden generator so that every side sees byte-identical inputs from a seed.
"""
from __future__ import annotations

import copy
import math
from typing import Dict, List

import numpy as np
import torch

# model.args of opencood/hypes_yaml/opv2v/lidar_only_with_noise/coalign/pointpillar_coalign.yaml:94-126
# (+ grid_size as injected by yaml_utils.load_point_pillar_params, yaml_utils.py:113-119)
OPV2V_ARGS = {
    "voxel_size": [0.4, 0.4, 4],
    "lidar_range": [-140.8, -40, -3, 140.8, 40, 1],
    "anchor_number": 2,
    "pillar_vfe": {"use_norm": True, "with_distance": False, "use_absolute_xyz": True, "num_filters": [64]},
    "point_pillar_scatter": {"num_features": 64},
    "base_bev_backbone": {"layer_nums": [3, 5, 8], "layer_strides": [2, 2, 2], "num_filters": [64, 128, 256],
                          "upsample_strides": [1, 2, 4], "num_upsample_filter": [128, 128, 128]},
    "fusion_method": "att",
    "att": {"feat_dim": [64, 128, 256]},
    "shrink_header": {"kernal_size": [3], "stride": [1], "padding": [1], "dim": [256], "input_dim": 384},
    "dir_args": {"dir_offset": 0.7853, "num_bins": 2, "anchor_yaw": [0, 90]},
}
# preprocess section of the same yaml (:47-56)
OPV2V_PRE = {"max_points_per_voxel": 32, "max_voxel_train": 32000, "max_voxel_test": 70000, "max_cav": 5}


def make_args(lidar_range=None, voxel_size=None, base=None):
    """Copy of the yaml model.args with `grid_size` injected like load_point_pillar_params does."""
    a = copy.deepcopy(base or OPV2V_ARGS)
    if lidar_range is not None:
        a["lidar_range"] = list(lidar_range)
    if voxel_size is not None:
        a["voxel_size"] = list(voxel_size)
    r, v = a["lidar_range"], a["voxel_size"]
    g = (np.array(r[3:6]) - np.array(r[0:3])) / np.array(v)
    a["point_pillar_scatter"]["grid_size"] = np.round(g).astype(np.int64)
    return a


def opv2v_args():
    return make_args()


def single_args(lidar_range=None, voxel_size=None):
    """model.args of opv2v/lidar_only_with_noise/pointpillar_single.yaml:69-95 (core_method point_pillar: BaseBEVBackbone,
    no fusion keys)."""
    a = make_args(lidar_range, voxel_size)
    a.pop("fusion_method", None)
    a.pop("att", None)
    return a


def uncertainty_args(lidar_range=None, voxel_size=None, uncertainty_dim=3):
    """model.args of opv2v/lidar_only_with_noise/coalign/pointpillar_uncertainty.yaml:73-93 (core_method
    point_pillar_uncertainty: the stage-1 detector whose boxes + uncertainties feed the pose-graph alignment): no shrink
    header, heads on the 384-channel decoded map, extra `unc_head`."""
    a = single_args(lidar_range, voxel_size)
    a.pop("shrink_header", None)
    a["uncertainty_dim"] = int(uncertainty_dim)
    return a


def dairv2x_args():
    """dairv2x/lidar_only_with_noise/coalign/pointpillar_coalign.yaml:52,57."""
    return make_args([-100.8, -40, -3.5, 100.8, 40, 1.5], [0.4, 0.4, 5])


# ------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------
def _bn(sd, prefix, c, g):
    sd[prefix + ".weight"] = torch.empty(c).uniform_(0.6, 1.4, generator=g)
    sd[prefix + ".bias"] = torch.empty(c).normal_(0.0, 0.1, generator=g)
    sd[prefix + ".running_mean"] = torch.empty(c).normal_(0.0, 0.2, generator=g)
    sd[prefix + ".running_var"] = torch.empty(c).uniform_(0.5, 1.5, generator=g)
    sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def _conv(shape, g, fan):
    return torch.empty(*shape).normal_(0.0, math.sqrt(2.0 / fan), generator=g)


def random_state_dict(args, seed=0, backbone="resnet") -> Dict[str, torch.Tensor]:
    """Random weights with the reference's state_dict key names and shapes, with
    non-trivial BatchNorm statistics so that BN-folding mistakes show up.  backbone="plain": the key layout of
    BaseBEVBackbone (opencood/models/sub_modules/base_bev_backbone.py:37-56) used by the single-agent `point_pillar`."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    nf = args["pillar_vfe"]["num_filters"][0]
    sd["pillar_vfe.pfn_layers.0.linear.weight"] = torch.empty(nf, 10).normal_(0.0, 0.3, generator=g)
    _bn(sd, "pillar_vfe.pfn_layers.0.norm", nf, g)
    bb = args["base_bev_backbone"]
    inpl = bb.get("inplanes", 64)
    plain = backbone == "plain"
    for li, (nb, st, pl) in enumerate(zip(bb["layer_nums"], bb["layer_strides"], bb["num_filters"])):
        if plain:           # BaseBEVBackbone.blocks[li] = [ZeroPad2d, Conv, BN, ReLU] + nb * [Conv, BN, ReLU]
            for j in range(nb + 1):
                cin = inpl if j == 0 else pl
                sd[f"backbone.blocks.{li}.{1 + 3 * j}.weight"] = _conv((pl, cin, 3, 3), g, 9 * cin)
                _bn(sd, f"backbone.blocks.{li}.{2 + 3 * j}", pl, g)
            inpl = pl
            continue
        for k in range(nb):
            p = f"backbone.resnet.layer{li}.{k}"
            cin = inpl if k == 0 else pl
            sd[p + ".conv1.weight"] = _conv((pl, cin, 3, 3), g, 9 * cin)
            _bn(sd, p + ".bn1", pl, g)
            sd[p + ".conv2.weight"] = _conv((pl, pl, 3, 3), g, 9 * pl * 2)
            _bn(sd, p + ".bn2", pl, g)
            if k == 0 and (st != 1 or inpl != pl):
                sd[p + ".downsample.0.weight"] = _conv((pl, cin, 1, 1), g, cin * 2)
                _bn(sd, p + ".downsample.1", pl, g)
        inpl = pl
    for i, (s, cu) in enumerate(zip(bb["upsample_strides"], bb["num_upsample_filter"])):
        ci = bb["num_filters"][i]
        sd[f"backbone.deblocks.{i}.0.weight"] = _conv((ci, cu, s, s), g, ci)
        _bn(sd, f"backbone.deblocks.{i}.1", cu, g)
    out_c = sum(bb["num_upsample_filter"])
    if "shrink_header" in args:
        sh = args["shrink_header"]
        cin = sh["input_dim"]
        for li, (ks, dim) in enumerate(zip(sh["kernal_size"], sh["dim"])):
            p = f"shrink_conv.layers.{li}.double_conv"
            sd[p + ".0.weight"] = _conv((dim, cin, ks, ks), g, ks * ks * cin)
            sd[p + ".0.bias"] = torch.empty(dim).normal_(0.0, 0.1, generator=g)
            sd[p + ".2.weight"] = _conv((dim, dim, 3, 3), g, 9 * dim)
            sd[p + ".2.bias"] = torch.empty(dim).normal_(0.0, 0.1, generator=g)
            cin = dim
        out_c = sh["dim"][-1]
    an = args["anchor_number"]
    heads = [("cls_head", an), ("reg_head", 7 * an)]
    if "uncertainty_dim" in args:                                   # point_pillar_uncertainty.py:34-35
        heads.append(("unc_head", args["uncertainty_dim"] * an))
    if "dir_args" in args:
        heads.append(("dir_head", args["dir_args"]["num_bins"] * an))
    for name, co in heads:
        sd[name + ".weight"] = _conv((co, out_c, 1, 1), g, out_c)
        sd[name + ".bias"] = torch.empty(co).normal_(0.0, 0.1, generator=g)
    return sd


# ------------------------------------------------------------------------------------------
# point clouds and poses
# ------------------------------------------------------------------------------------------
def lidar_cloud(rng: np.random.Generator, n_points: int, lidar_range, sigma=35.0) -> np.ndarray:
    """LiDAR-like cloud in the agent's own frame: r = |N(0,sigma)|+2, azimuth U(0,2pi),
    z uniform strictly inside the z range, intensity U(0,1).  float32 (P,4)."""
    r = np.abs(rng.normal(0.0, sigma, n_points)) + 2.0
    az = rng.uniform(0.0, 2 * np.pi, n_points)
    zlo, zhi = lidar_range[2], lidar_range[5]
    z = rng.uniform(zlo + 0.025 * (zhi - zlo), zhi - 0.1 * (zhi - zlo), n_points)
    pts = np.stack([r * np.cos(az), r * np.sin(az), z, rng.uniform(0, 1, n_points)], axis=1)
    return pts.astype(np.float32)


def _pose_matrix(pose):
    x, y, z, roll, yaw, pitch = [float(v) for v in pose]
    cy, sy = math.cos(math.radians(yaw)), math.sin(math.radians(yaw))
    cr, sr = math.cos(math.radians(roll)), math.sin(math.radians(roll))
    cp, sp = math.cos(math.radians(pitch)), math.sin(math.radians(pitch))
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    ry = np.array([[cp, 0, -sp], [0, 1.0, 0], [sp, 0, cp]])
    rx = np.array([[1.0, 0, 0], [0, cr, sr], [0, -sr, cr]])
    m = np.identity(4)
    m[:3, :3] = rz @ ry @ rx
    m[:3, 3] = (x, y, z)
    return m


def make_scene(seed: int, n_agents: int, n_points: int, lidar_range, max_cav=5, pose_noise=False,
               spread=50.0, sigma=35.0):
    """One scene: per-agent clouds (own frame, proj_first=false) + pairwise_t_matrix (L,L,4,4) f64
    with [i,j] = T_j^-1 T_i (transformation_utils.py:22-67), identity padded."""
    rng = np.random.default_rng(seed)
    clouds = [lidar_cloud(rng, n_points, lidar_range, sigma) for _ in range(n_agents)]
    poses = [np.zeros(6)]
    for _ in range(1, n_agents):
        poses.append(np.array([rng.uniform(-spread, spread), rng.uniform(-spread, spread), 0, 0,
                               rng.uniform(-180, 180), 0]))
    if pose_noise:      # pose_utils.py:43-73 noise model, N(0,0.2 m) / N(0,0.2 deg)
        poses = [p + np.array([rng.normal(0, 0.2), rng.normal(0, 0.2), 0, 0, rng.normal(0, 0.2), 0])
                 for p in poses]
    ts = [_pose_matrix(p) for p in poses]
    pw = np.tile(np.eye(4), (max_cav, max_cav, 1, 1))
    for i in range(n_agents):
        for j in range(n_agents):
            if i != j:
                pw[i, j] = np.linalg.solve(ts[j], ts[i])
    return {"points": clouds, "poses": poses, "pairwise_t_matrix": pw}


def post_params(lidar_range=None, voxel_size=None, H_map=None, W_map=None):
    """`postprocess` block of opv2v/lidar_only_with_noise/coalign/pointpillar_coalign.yaml:69-90 plus the
    anchor_args fields that yaml_utils.load_point_pillar_params (yaml_utils.py:121-135) derives.  With H_map/W_map the
    lidar range is chosen so that the head maps are H_map x W_map (feature_stride 2, 0.4 m voxels)."""
    import math
    vs = list(voxel_size or [0.4, 0.4, 4])
    if lidar_range is None:
        if H_map is None:
            lidar_range = [-140.8, -40, -3, 140.8, 40, 1]
        else:
            lidar_range = [-W_map * vs[0], -H_map * vs[1], -3, W_map * vs[0], H_map * vs[1], 1]
    r = list(lidar_range)
    return {"core_method": "VoxelPostprocessor", "gt_range": r,
            "anchor_args": {"cav_lidar_range": r, "l": 3.9, "w": 1.6, "h": 1.56, "r": [0, 90], "feature_stride": 2, "num": 2,
                            "vw": vs[0], "vh": vs[1], "vd": vs[2],
                            "W": math.ceil((r[3] - r[0]) / vs[0]), "H": math.ceil((r[4] - r[1]) / vs[1]),
                            "D": math.ceil((r[5] - r[2]) / vs[2])},
            "target_args": {"pos_threshold": 0.6, "neg_threshold": 0.45, "score_threshold": 0.20},
            "order": "hwl", "max_num": 100, "nms_thresh": 0.15,
            "dir_args": {"dir_offset": 0.7853, "num_bins": 2, "anchor_yaw": [0, 90]}}


# ------------------------------------------------------------------------------------------
# loss
# ------------------------------------------------------------------------------------------
def loss_args():
    """yaml `loss.args` of opv2v/lidar_only_with_noise/coalign/pointpillar_coalign.yaml:132-147."""
    return {"pos_cls_weight": 2.0,
            "cls": {"type": "SigmoidFocalLoss", "alpha": 0.25, "gamma": 2.0, "weight": 2.0},
            "reg": {"type": "WeightedSmoothL1Loss", "sigma": 3.0, "codewise": True, "weight": 2.0},
            "dir": {"type": "WeightedSoftmaxClassificationLoss", "weight": 0.2,
                    "args": {"dir_offset": 0.7853, "num_bins": 2, "anchor_yaw": [0, 90]}}}


def loss_case(seed, n=2, H=12, W=20, A=2, n_pos=9, empty_samples=(), dtype=np.float64):
    """Synthetic head outputs + label tensors with the structure `VoxelPostprocessor.generate_label` / `collate_batch`
    produce (voxel_postprocessor.py:84-241): float64 (n,H,W,A) pos / neg masks (positives sparse, a don't-care band
    around them, everything else negative), float64 (n,H,W,7A) regression targets that are non-zero on positives only."""
    rng = np.random.default_rng(seed)
    cls = rng.normal(-2.0, 2.0, (n, A, H, W)).astype(np.float32)
    reg = rng.normal(0.0, 0.4, (n, 7 * A, H, W)).astype(np.float32)
    dr = rng.normal(0.0, 1.5, (n, 2 * A, H, W)).astype(np.float32)
    pos = np.zeros((n, H, W, A), dtype)
    neg = np.ones((n, H, W, A), dtype)
    tgt = np.zeros((n, H, W, 7 * A), dtype)
    for b in range(n):
        if b in empty_samples:
            continue
        for _ in range(n_pos):
            h, w, a = int(rng.integers(1, H - 1)), int(rng.integers(1, W - 1)), int(rng.integers(0, A))
            neg[b, h - 1:h + 2, w - 1:w + 2, :] = 0                       # don't-care band
            pos[b, h, w, a] = 1
            t = rng.normal(0.0, 0.3, 7)
            t[6] = rng.uniform(-3.0, 3.0)                                   # yaw residual: exercises both direction bins
            tgt[b, h, w, 7 * a:7 * a + 7] = t
            cls[b, a, h, w] += 3.0
    # a few large regression errors (linear branch of the smooth L1) and exact zeros (|diff| = 0)
    reg[:, :, 0, 0] = 2.5
    reg[:, :, 1, 1] = 0.0
    return {"cls": cls, "reg": reg, "dir": dr, "pos": pos, "neg": neg, "tgt": tgt}


# ------------------------------------------------------------------------------------------
# camera BEV half (BASELINE configs[4]): BevEncodeMSFusion on the splat output
# ------------------------------------------------------------------------------------------
def random_camera_bev_state_dict(seed=0, in_channels=128) -> Dict[str, torch.Tensor]:
    """Random-init weights with the key names / shapes of the reference's BevEncodeMSFusion (lss_submodule.py:357-388:
    7x7/s2 stem, resnet18 layer1-3, two Up blocks, down_layer); BatchNorm statistics randomised like random_state_dict."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd: Dict[str, torch.Tensor] = {}

    def bn(prefix, c):
        sd[prefix + ".weight"] = torch.rand(c, generator=g) * 0.8 + 0.6
        sd[prefix + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[prefix + ".running_var"] = torch.rand(c, generator=g) * 0.8 + 0.6
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)

    def conv(name, cout, cin, k):
        sd[name] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))

    conv("conv1.weight", 64, in_channels, 7)
    bn("bn1", 64)
    cin = 64
    for li, c in ((1, 64), (2, 128), (3, 256)):
        for b in range(2):
            p = f"layer{li}.{b}"
            conv(p + ".conv1.weight", c, cin if b == 0 else c, 3)
            bn(p + ".bn1", c)
            conv(p + ".conv2.weight", c, c, 3)
            bn(p + ".bn2", c)
            if b == 0 and cin != c:
                conv(p + ".downsample.0.weight", c, cin, 1)
                bn(p + ".downsample.1", c)
        cin = c
    for name, ci in (("up_layer2", 128 + 256), ("up_layer1", 64 + 256)):
        conv(name + ".conv.0.weight", 256, ci, 3)
        bn(name + ".conv.1", 256)
        conv(name + ".conv.3.weight", 256, 256, 3)
        bn(name + ".conv.4", 256)
    conv("down_layer.0.weight", 256, 256, 3)
    sd["down_layer.0.bias"] = torch.randn(256, generator=g) * 0.05
    conv("down_layer.2.weight", 128, 256, 3)
    sd["down_layer.2.bias"] = torch.randn(128, generator=g) * 0.05
    return sd


def camera_bev_case(record_len, seed, hw=240, in_channels=128, max_cav=5, voxel=0.4):
    """Synthetic splat output (sumN, 128, hw, hw) float32 (sparse-ish, non-negative like a pooled frustum feature map) and
    pairwise_t_matrix (B, L, L, 4, 4) float64 for agents spread inside the [-hw*voxel/2, hw*voxel/2] grid."""
    rng = np.random.default_rng(5000 + seed)
    n = int(sum(record_len))
    x = rng.standard_normal((n, in_channels, hw, hw)).astype(np.float32)
    x *= (rng.random((n, 1, hw, hw)) < 0.6).astype(np.float32)
    half = hw * voxel / 2
    pws = []
    for k in record_len:
        poses = [[0, 0, 0, 0, 0, 0]] + [[float(rng.uniform(-half / 2, half / 2)), float(rng.uniform(-half / 2, half / 2)), 0, 0,
                                         float(rng.uniform(-180, 180)), 0] for _ in range(k - 1)]
        ts = [_pose_matrix(p) for p in poses]
        pw = np.tile(np.eye(4), (max_cav, max_cav, 1, 1))
        for i in range(k):
            for j in range(k):
                if i != j:
                    pw[i, j] = np.linalg.solve(ts[j], ts[i])
        pws.append(pw)
    return x, np.stack(pws)


def lift_splat_case(seed=0, B=2, N=4, C=64, final_dim=(96, 128), downsample=8, n_bins=16):
    """Small lift + splat problem with the structure of the camera yaml (lss_coalign_fusion.yaml:28-43): N pinhole cameras
    looking outwards from the agent, LID depth bins, a square BEV grid.  Returns numpy arrays + the config dicts."""
    rng = np.random.default_rng(7000 + seed)
    grid_conf = {"xbound": [-16.0, 16.0, 0.4], "ybound": [-16.0, 16.0, 0.4], "zbound": [-10.0, 10.0, 20.0],
                 "ddiscr": [2, 18, n_bins], "mode": "LID"}
    fH, fW = final_dim[0] // downsample, final_dim[1] // downsample
    rots = np.zeros((B, N, 3, 3), np.float32)
    trans = np.zeros((B, N, 3), np.float32)
    intr = np.zeros((B, N, 3, 3), np.float32)
    post_rots = np.zeros((B, N, 3, 3), np.float32)
    post_trans = np.zeros((B, N, 3), np.float32)
    cam_to_ego = np.array([[0, 0, 1.0], [-1.0, 0, 0], [0, -1.0, 0]])          # camera (x right, y down, z fwd) -> ego (x fwd, y left, z up)
    for b in range(B):
        for n in range(N):
            yaw = np.deg2rad(90.0 * n + rng.uniform(-10, 10))
            rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
            rots[b, n] = (rz @ cam_to_ego).astype(np.float32)
            trans[b, n] = [rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(1.2, 1.8)]
            f = rng.uniform(0.9, 1.1) * final_dim[1] / 2
            intr[b, n] = [[f, 0, final_dim[1] / 2], [0, f, final_dim[0] / 2], [0, 0, 1]]
            sc = rng.uniform(0.8, 0.85)
            post_rots[b, n] = np.diag([sc, sc, 1.0]).astype(np.float32)
            post_trans[b, n] = [rng.uniform(-8, 0), rng.uniform(-8, 0), 0]
    return {"grid_conf": grid_conf, "final_dim": list(final_dim), "downsample": downsample,
            "depth_logit": rng.standard_normal((B * N, n_bins, fH, fW)).astype(np.float32) * 2.0,
            "x_img": rng.standard_normal((B * N, C, fH, fW)).astype(np.float32),
            "rots": rots, "trans": trans, "intrins": intr, "post_rots": post_rots, "post_trans": post_trans}
