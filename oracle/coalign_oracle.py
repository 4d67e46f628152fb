"""CPU restatement (torch fp32/fp64 on host) of the CoAlign forward, stages A3..A13.

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.  Every function cites the reference
lines it follows (paths relative to /root/reference).  The restatement is *functional*: it takes
the reference ``state_dict`` (same key names) and the yaml ``model.args`` dict, so it can be
checked against the reference module with identical weights (tests/golden/gen_golden.py).

`forward` is the inference path (eval-mode BatchNorm, no autograd).  `forward_train` (bottom of the file) is the same graph
with train-mode BatchNorm (batch statistics) and autograd enabled: the oracle of the training step,
pinned by tests/golden/train_small.npz (losses, gradients and running-stat updates of the unmodified reference in
`.train()` mode).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def bn_affine(sd: Dict[str, torch.Tensor], prefix: str, eps: float):
    """Eval-mode BatchNorm as y = x*scale + shift.

    eps: 1e-3 for PFN (opencood/models/sub_modules/pillar_vfe.py:25) and deblocks
    (base_bev_backbone_resnet.py:62-63); 1e-5 (nn.BatchNorm2d default) inside BasicBlocks
    and downsample branches (resblock.py:38-39,192-196).
    """
    g = sd[prefix + ".weight"].double()
    b = sd[prefix + ".bias"].double()
    m = sd[prefix + ".running_mean"].double()
    v = sd[prefix + ".running_var"].double()
    scale = g / torch.sqrt(v + eps)
    shift = b - m * scale
    return scale.float(), shift.float()


# train-mode switch of the BatchNorm helpers (set by forward_train only): batch statistics instead of the running ones;
# the batch mean / unbiased variance of every layer are recorded so that the running-stat update can be checked
_TRAIN = {"on": False, "batch_stats": None}


def _record_stats(prefix, x, dims):
    if _TRAIN["batch_stats"] is not None:
        with torch.no_grad():
            _TRAIN["batch_stats"][prefix] = (x.mean(dims), x.var(dims, unbiased=True))


def _bn2d(x, sd, prefix, eps):
    if _TRAIN["on"]:                                    # nn.BatchNorm2d.forward in training mode
        _record_stats(prefix, x, (0, 2, 3))
        return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.0, eps)
    # F.batch_norm in eval mode: (x-mean)/sqrt(var+eps)*w+b, exactly what nn.BatchNorm2d does.
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, eps)


# --------------------------------------------------------------------------------------
# A3 + A4: PillarVFE / PFNLayer          opencood/models/sub_modules/pillar_vfe.py:31-53,105-155
# --------------------------------------------------------------------------------------
def pillar_vfe(sd, args, voxel_features, voxel_coords, voxel_num_points):
    """(M,32,4),(M,4)[a,z,y,x],(M,) -> (M,64) pillar features."""
    vx, vy, vz = [float(v) for v in args["voxel_size"]]
    rng = [float(v) for v in args["lidar_range"]]
    x_off, y_off, z_off = vx / 2 + rng[0], vy / 2 + rng[1], vz / 2 + rng[2]   # pillar_vfe.py:87-89
    vf = voxel_features.to(sd["pillar_vfe.pfn_layers.0.linear.weight"].dtype)   # float32 (float64 state_dict: exactness checks)
    n = voxel_num_points.to(vf.dtype).view(-1, 1, 1)
    mean = vf[:, :, :3].sum(dim=1, keepdim=True) / n                           # :118-120
    f_cluster = vf[:, :, :3] - mean                                             # :121
    cf = voxel_coords.to(vf.dtype)
    f_center = torch.zeros_like(vf[:, :, :3])                                   # :123-132
    f_center[:, :, 0] = vf[:, :, 0] - (cf[:, 3].unsqueeze(1) * vx + x_off)
    f_center[:, :, 1] = vf[:, :, 1] - (cf[:, 2].unsqueeze(1) * vy + y_off)
    f_center[:, :, 2] = vf[:, :, 2] - (cf[:, 1].unsqueeze(1) * vz + z_off)
    feats = torch.cat([vf, f_cluster, f_center], dim=-1)                        # use_absolute_xyz :134-137
    slots = torch.arange(vf.shape[1], dtype=torch.int32).view(1, -1)
    mask = (voxel_num_points.int().view(-1, 1) > slots).unsqueeze(-1).to(vf.dtype)  # :145-149
    feats = feats * mask
    w = sd["pillar_vfe.pfn_layers.0.linear.weight"]                             # (64,10) no bias, :24
    x = feats @ w.t()                                                           # :31-40
    if _TRAIN["on"]:                                    # BatchNorm1d(64) over (M, 64, 32): padded slots take part (:41-44)
        pre = "pillar_vfe.pfn_layers.0.norm"
        _record_stats(pre, x, (0, 1))
        x = F.batch_norm(x.permute(0, 2, 1), None, None, sd[pre + ".weight"], sd[pre + ".bias"], True, 0.0,
                         1e-3).permute(0, 2, 1)
    else:
        scale, shift = bn_affine(sd, "pillar_vfe.pfn_layers.0.norm", 1e-3)      # :25,41-44
        x = x * scale + shift
    x = F.relu(x)                                                               # :45
    return x.max(dim=1)[0]                                                      # :46 (padded slots included)


# --------------------------------------------------------------------------------------
# A5: PointPillarScatter                 opencood/models/sub_modules/point_pillar_scatter.py:15-72
# --------------------------------------------------------------------------------------
def scatter(pillar_features, voxel_coords, n_agents, ny, nx):
    c = pillar_features.shape[1]
    canvas = torch.zeros(n_agents, c, ny * nx, dtype=pillar_features.dtype)
    a = voxel_coords[:, 0].long()
    idx = (voxel_coords[:, 1] + voxel_coords[:, 2] * nx + voxel_coords[:, 3]).long()   # :54
    canvas[a, :, idx] = pillar_features                                                # :61
    return canvas.view(n_agents, c, ny, nx)


# --------------------------------------------------------------------------------------
# A6: ResNetModified / BasicBlock        opencood/models/sub_modules/resblock.py:53-69,177-221
# --------------------------------------------------------------------------------------
def basic_block(sd, prefix, x, stride, has_ds):
    out = F.conv2d(x, sd[prefix + ".conv1.weight"], None, stride, 1)            # conv3x3 :12-15
    out = F.relu(_bn2d(out, sd, prefix + ".bn1", 1e-5))
    out = F.conv2d(out, sd[prefix + ".conv2.weight"], None, 1, 1)
    out = _bn2d(out, sd, prefix + ".bn2", 1e-5)
    if has_ds:                                                                   # :192-196
        idn = F.conv2d(x, sd[prefix + ".downsample.0.weight"], None, stride, 0)
        idn = _bn2d(idn, sd, prefix + ".downsample.1", 1e-5)
    else:
        idn = x
    return F.relu(out + idn)                                                     # :66-67


def encoder(sd, args, x) -> List[torch.Tensor]:
    """base_bev_backbone_resnet.py:114-119 -> resblock.py:212-221."""
    bb = args["base_bev_backbone"]
    inplanes = bb.get("inplanes", 64)
    feats = []
    for li, (nblk, stride, planes) in enumerate(zip(bb["layer_nums"], bb["layer_strides"], bb["num_filters"])):
        for k in range(nblk):
            s = stride if k == 0 else 1
            has_ds = (k == 0) and (s != 1 or inplanes != planes)                 # resblock.py:191
            x = basic_block(sd, f"backbone.resnet.layer{li}.{k}", x, s, has_ds)
        inplanes = planes
        feats.append(x)
    return feats


# --------------------------------------------------------------------------------------
# A10: normalize_pairwise_tfm            opencood/utils/transformation_utils.py:69-91
# --------------------------------------------------------------------------------------
def normalize_pairwise_tfm(pairwise_t_matrix, H, W, discrete_ratio, downsample_rate=1):
    t = pairwise_t_matrix.double()
    a = torch.zeros(*t.shape[:3], 2, 3, dtype=torch.float64)
    a[..., 0, 0] = t[..., 0, 0]
    a[..., 0, 1] = t[..., 0, 1] * H / W
    a[..., 0, 2] = t[..., 0, 3] / (downsample_rate * discrete_ratio * W) * 2
    a[..., 1, 0] = t[..., 1, 0] * W / H
    a[..., 1, 1] = t[..., 1, 1]
    a[..., 1, 2] = t[..., 1, 3] / (downsample_rate * discrete_ratio * H) * 2
    return a


# --------------------------------------------------------------------------------------
# A11: warp_affine_simple                opencood/models/sub_modules/torch_transformation_utils.py:322-331
# (F.affine_grid + F.grid_sample restated explicitly: bilinear, zeros padding, align_corners=False)
# --------------------------------------------------------------------------------------
def warp_affine_simple(src, M):
    """src (N,C,H,W) f32, M (N,2,3) f64 -> (N,C,H,W)."""
    N, C, H, W = src.shape
    M = M.double()
    xs = (2.0 * torch.arange(W, dtype=torch.float64) + 1.0) / W - 1.0           # base grid, align_corners=False
    ys = (2.0 * torch.arange(H, dtype=torch.float64) + 1.0) / H - 1.0
    gx = (M[:, 0, 0].view(N, 1, 1) * xs.view(1, 1, W) + M[:, 0, 1].view(N, 1, 1) * ys.view(1, H, 1)
          + M[:, 0, 2].view(N, 1, 1))
    gy = (M[:, 1, 0].view(N, 1, 1) * xs.view(1, 1, W) + M[:, 1, 1].view(N, 1, 1) * ys.view(1, H, 1)
          + M[:, 1, 2].view(N, 1, 1))
    gx = gx.to(src.dtype)                                                        # `.to(src)` :330
    gy = gy.to(src.dtype)
    ix = ((gx + 1) * W - 1) / 2                                                  # grid_sample unnormalise
    iy = ((gy + 1) * H - 1) / 2
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    wx1 = ix - x0
    wy1 = iy - y0
    wx0 = 1 - wx1
    wy0 = 1 - wy1
    out = torch.zeros_like(src)
    flat = src.reshape(N, C, H * W)
    for dy, dx, wgt in ((0, 0, wy0 * wx0), (0, 1, wy0 * wx1), (1, 0, wy1 * wx0), (1, 1, wy1 * wx1)):
        xi = (x0 + dx).long()
        yi = (y0 + dy).long()
        valid = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        lin = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).view(N, 1, H * W).expand(N, C, H * W)
        v = torch.gather(flat, 2, lin).view(N, C, H, W)
        out = out + v * (wgt * valid.to(src.dtype)).view(N, 1, H, W)
    return out


# --------------------------------------------------------------------------------------
# A12/A13: AttFusion (ego row)           opencood/models/fuse_modules/fusion_in_one.py:96-136
#          ScaledDotProductAttention      opencood/models/fuse_modules/att_fuse.py:43-47
# --------------------------------------------------------------------------------------
def att_fusion(x, record_len, affine, method="att"):
    """x (sumN,C,H,W); record_len (B,); affine (B,L,L,2,3) f64 -> (B,C,H,W)."""
    C = x.shape[1]
    out = []
    start = 0
    for b, n in enumerate([int(v) for v in record_len]):                         # regroup :21-24
        xb = x[start:start + n]
        start += n
        w = warp_affine_simple(xb, affine[b, 0, :n])                             # :125-128 (ego row i=0)
        if method == "max":                                                      # MaxFusion :83-86
            out.append(w.max(dim=0)[0])
            continue
        score = (w[0:1] * w).sum(dim=1) / np.sqrt(C)                             # (n,H,W)   att_fuse.py:44
        attn = torch.softmax(score, dim=0)                                       # :45
        out.append((attn.unsqueeze(1) * w).sum(dim=0))                           # :46, row 0 kept (:132)
    return torch.stack(out)


# --------------------------------------------------------------------------------------
# A7: deblocks + concat                  opencood/models/sub_modules/base_bev_backbone_resnet.py:52-65,121-138
# --------------------------------------------------------------------------------------
def decoder(sd, args, fused: List[torch.Tensor]):
    bb = args["base_bev_backbone"]
    ups = []
    for i, s in enumerate(bb["upsample_strides"]):
        y = F.conv_transpose2d(fused[i], sd[f"backbone.deblocks.{i}.0.weight"], None, stride=s)
        y = F.relu(_bn2d(y, sd, f"backbone.deblocks.{i}.1", 1e-3))
        ups.append(y)
    return torch.cat(ups, dim=1)


# --------------------------------------------------------------------------------------
# A8: shrink header                      opencood/models/sub_modules/downsample_conv.py:7-50
# --------------------------------------------------------------------------------------
def shrink(sd, args, x):
    sh = args["shrink_header"]
    for li, (k, s, p) in enumerate(zip(sh["kernal_size"], sh["stride"], sh["padding"])):
        pre = f"shrink_conv.layers.{li}.double_conv"
        x = F.relu(F.conv2d(x, sd[pre + ".0.weight"], sd[pre + ".0.bias"], s, p))
        x = F.relu(F.conv2d(x, sd[pre + ".2.weight"], sd[pre + ".2.bias"], 1, 1))
    return x


# --------------------------------------------------------------------------------------
# A9: heads                              opencood/models/point_pillar_baseline_multiscale.py:55-63,126-133
# --------------------------------------------------------------------------------------
def heads(sd, x):
    out = {"cls_preds": F.conv2d(x, sd["cls_head.weight"], sd["cls_head.bias"]),
           "reg_preds": F.conv2d(x, sd["reg_head.weight"], sd["reg_head.bias"])}
    if "unc_head.weight" in sd:                                                  # point_pillar_uncertainty.py:66,70
        out["unc_preds"] = F.conv2d(x, sd["unc_head.weight"], sd["unc_head.bias"])
    if "dir_head.weight" in sd:
        out["dir_preds"] = F.conv2d(x, sd["dir_head.weight"], sd["dir_head.bias"])
    return out


# --------------------------------------------------------------------------------------
# whole forward                          opencood/models/point_pillar_baseline_multiscale.py:93-135
# --------------------------------------------------------------------------------------
@torch.no_grad()
def forward(sd, args, data_dict, stages: Optional[dict] = None):
    return _forward(sd, args, data_dict, stages)


def forward_train(sd, args, data_dict, stages: Optional[dict] = None):
    """The reference forward in `.train()` mode (train.py:109-114): BatchNorm layers use batch statistics, autograd is on
    (pass a state_dict whose parameters require grad).  Returns (output dict, {bn prefix: (batch mean, unbiased batch
    var)}); the running statistics a step would leave behind are  (1-m)*running + m*batch  with m = 0.1 inside the
    BasicBlocks (nn.BatchNorm2d default, resblock.py:38-39) and 0.01 for the PFN norm and the deblocks
    (pillar_vfe.py:25, base_bev_backbone_resnet.py:62-63)."""
    _TRAIN["on"], _TRAIN["batch_stats"] = True, {}
    try:
        out = _forward(sd, args, data_dict, stages)
        return out, _TRAIN["batch_stats"]
    finally:
        _TRAIN["on"], _TRAIN["batch_stats"] = False, None


def _forward(sd, args, data_dict, stages: Optional[dict] = None):
    pl = data_dict["processed_lidar"]
    vf, vc, vn = pl["voxel_features"], pl["voxel_coords"], pl["voxel_num_points"]
    record_len = data_dict["record_len"]
    nx, ny, nz = [int(v) for v in args["point_pillar_scatter"]["grid_size"]]
    n_agents = int(sum(int(v) for v in record_len))
    pf = pillar_vfe(sd, args, vf, vc, vn)                                        # :104
    canvas = scatter(pf, vc, n_agents, ny, nx)                                   # :106
    affine = normalize_pairwise_tfm(data_dict["pairwise_t_matrix"], ny, nx,
                                    float(args["voxel_size"][0]))                # :108-109
    feats = encoder(sd, args, canvas)                                            # :117
    fused = [att_fusion(f, record_len, affine, args.get("fusion_method", "att")) for f in feats]  # :119-120
    dec = decoder(sd, args, fused)                                               # :121
    sh = shrink(sd, args, dec) if "shrink_header" in args else dec               # :123-124
    out = heads(sd, sh)                                                          # :126-133
    if stages is not None:
        stages.update(pillar_features=pf, canvas=canvas, affine=affine, feats=feats, fused=fused,
                      decoded=dec, shrunk=sh)
    return out


# --------------------------------------------------------------------------------------
# single-agent PointPillar (BASELINE configs[0]): opencood/models/point_pillar.py:52-84 with
# BaseBEVBackbone                          opencood/models/sub_modules/base_bev_backbone.py:37-56,96-125
# --------------------------------------------------------------------------------------
def plain_encoder(sd, args, x) -> List[torch.Tensor]:
    """BaseBEVBackbone.blocks: [ZeroPad2d(1), Conv3x3(stride s, pad 0), BN(eps 1e-3), ReLU] + n x [Conv3x3(pad 1), BN, ReLU]."""
    bb = args["base_bev_backbone"]
    feats = []
    for li, (nb, stride) in enumerate(zip(bb["layer_nums"], bb["layer_strides"])):
        for j in range(nb + 1):
            pre = f"backbone.blocks.{li}."
            if j == 0:
                x = F.conv2d(F.pad(x, (1, 1, 1, 1)), sd[pre + "1.weight"], None, stride, 0)
            else:
                x = F.conv2d(x, sd[pre + f"{1 + 3 * j}.weight"], None, 1, 1)
            x = F.relu(_bn2d(x, sd, pre + f"{2 + 3 * j}", 1e-3))
        feats.append(x)
    return feats


@torch.no_grad()
def forward_single(sd, args, data_dict, stages: Optional[dict] = None):
    """PointPillar.forward (point_pillar.py:52-84): every sample of the batch is an independent single-agent frame."""
    pl = data_dict["processed_lidar"]
    vf, vc, vn = pl["voxel_features"], pl["voxel_coords"], pl["voxel_num_points"]
    nx, ny, nz = [int(v) for v in args["point_pillar_scatter"]["grid_size"]]
    n = int(vc[:, 0].max()) + 1                                                  # point_pillar_scatter.py:41
    pf = pillar_vfe(sd, args, vf, vc, vn)
    canvas = scatter(pf, vc, n, ny, nx)
    resnet = args["base_bev_backbone"].get("resnet", False)                      # point_pillar.py:27-31
    feats = encoder(sd, args, canvas) if resnet else plain_encoder(sd, args, canvas)
    dec = decoder(sd, args, feats)                                               # base_bev_backbone.py:107-119
    sh = shrink(sd, args, dec) if "shrink_header" in args else dec
    out = heads(sd, sh)
    if stages is not None:
        stages.update(pillar_features=pf, canvas=canvas, feats=feats, decoded=dec, shrunk=sh)
    return out


# --------------------------------------------------------------------------------------
# synthetic inputs shared by tests / bench / golden generation
# --------------------------------------------------------------------------------------
def pose_to_tfm(pose):
    """[x,y,z,roll,yaw,pitch] (degrees) -> 4x4 world-from-agent matrix; same matrix as x_to_world
    (opencood/utils/transformation_utils.py:263-306), written as Rz(yaw)·Ry(-pitch)·Rx(-roll)."""
    x, y, z, roll, yaw, pitch = [float(v) for v in pose]

    def rot(axis, deg):
        c, s = math.cos(math.radians(deg)), math.sin(math.radians(deg))
        i, j = {"x": (1, 2), "y": (2, 0), "z": (0, 1)}[axis]
        r = np.identity(3)
        r[i, i], r[i, j], r[j, i], r[j, j] = c, -s, s, c
        return r

    m = np.identity(4)
    m[:3, :3] = rot("z", yaw) @ rot("y", -pitch) @ rot("x", -roll)
    m[:3, 3] = (x, y, z)
    return m


def pairwise_from_poses(poses, max_cav):
    """get_pairwise_transformation, proj_first=False (transformation_utils.py:22-67):
    pairwise[i,j] = T_j^-1 T_i, identity padded to (L,L,4,4) float64."""
    L = max_cav
    pw = np.tile(np.eye(4), (L, L, 1, 1))
    ts = [pose_to_tfm(p) for p in poses]
    for i in range(len(ts)):
        for j in range(len(ts)):
            if i != j:
                pw[i, j] = np.linalg.solve(ts[j], ts[i])
    return pw
