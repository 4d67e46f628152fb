"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement (torch fp32 on the host) of the BEV half of the camera
model of BASELINE configs[4] - `BevEncodeMSFusion.forward` (/root/reference/opencood/models/sub_modules/lss_submodule.py:
357-417) with `Up` (:19-38), torchvision's resnet18 BasicBlocks (layer1-3) and AttFusion / MaxFusion per scale - as a pure
function of the module's `state_dict` (same key names), eval-mode BatchNorm.  Pinned by tests/golden/camera_bev_small.npz,
generated from the unmodified reference class (tests/golden/gen_golden_camera.py)."""
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import coalign_oracle as O


def _bn(x, sd, prefix, eps=1e-5):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.0, eps)


def _basic_block(sd, p, x, stride):
    """torchvision.models.resnet.BasicBlock: conv3x3(stride)-bn-relu-conv3x3-bn (+ 1x1/stride downsample-bn) -> add -> relu."""
    out = F.relu(_bn(F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1), sd, p + ".bn1"))
    out = _bn(F.conv2d(out, sd[p + ".conv2.weight"], None, 1, 1), sd, p + ".bn2")
    idn = x
    if (p + ".downsample.0.weight") in sd:
        idn = _bn(F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0), sd, p + ".downsample.1")
    return F.relu(out + idn)


def _up(sd, p, x1, x2):
    """Up.forward (lss_submodule.py:35-38): bilinear x2 (align_corners=True), cat([x2, up(x1)]), 2 x (conv3x3 + BN + ReLU)."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    x = torch.cat([x2, x1], dim=1)
    x = F.relu(_bn(F.conv2d(x, sd[p + ".conv.0.weight"], None, 1, 1), sd, p + ".conv.1"))
    return F.relu(_bn(F.conv2d(x, sd[p + ".conv.3.weight"], None, 1, 1), sd, p + ".conv.4"))


def _down(sd, x):
    x = F.relu(F.conv2d(x, sd["down_layer.0.weight"], sd["down_layer.0.bias"], 1, 1))
    return F.relu(F.conv2d(x, sd["down_layer.2.weight"], sd["down_layer.2.bias"], 1, 1))


@torch.no_grad()
def bev_encode_ms_fusion(sd: Dict[str, torch.Tensor], x, record_len, pairwise_t_matrix, discrete_ratio: float,
                         method: str = "att", stages: Optional[dict] = None):
    """x (sumN, inC, H, W) splat output; returns (x_single (sumN,128,H/2,W/2), x_fuse (B,128,H/2,W/2))."""
    _, _C, H, W = x.shape
    affine = O.normalize_pairwise_tfm(pairwise_t_matrix, H, W, discrete_ratio, 1)            # :390
    x = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], None, 2, 3), sd, "bn1"))                  # :392-394
    feats = []
    for li, stride in ((1, 1), (2, 2), (3, 2)):                                              # :396-398
        x = _basic_block(sd, f"layer{li}.0", x, stride)
        x = _basic_block(sd, f"layer{li}.1", x, 1)
        feats.append(x)
    x1, x2, x3 = feats
    x_single = _down(sd, _up(sd, "up_layer1", _up(sd, "up_layer2", x3, x2), x1))             # :399
    fused = [O.att_fusion(f, record_len, affine, method) for f in feats]                     # :401-403
    x_fuse = _down(sd, _up(sd, "up_layer1", _up(sd, "up_layer2", fused[2], fused[1]), fused[0]))   # :405
    if stages is not None:
        stages.update(feats=feats, fused=fused)
    return x_single, x_fuse


# ----------------------------------------------------------------------------------------------------------------------
# lift + splat: /root/reference/opencood/models/lift_splat_shoot.py:64-169 (create_frustum, get_geometry, get_cam_feats'
# reshape, voxel_pooling) and the "lift" line of CamEncode.forward (lss_submodule.py:134-136: depth softmax (x) features);
# depth_discretization / gen_dx_bx from utils/camera_utils.py:129-135,186-195.
# ----------------------------------------------------------------------------------------------------------------------
import numpy as np      # noqa: E402


def depth_bins(d_min, d_max, num_bins, mode):
    if mode == "UD":
        return d_min + (d_max - d_min) / num_bins * np.arange(num_bins)
    if mode == "LID":
        bin_size = 2 * (d_max - d_min) / (num_bins * (1 + num_bins))
        return d_min + bin_size * (np.arange(num_bins) * np.arange(1, 1 + num_bins)) / 2
    raise NotImplementedError(mode)


def gen_dx_bx(xbound, ybound, zbound):
    dx = torch.Tensor([row[2] for row in [xbound, ybound, zbound]])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in [xbound, ybound, zbound]])
    nx = torch.LongTensor([(row[1] - row[0]) / row[2] for row in [xbound, ybound, zbound]])
    return dx, bx, nx


def frustum(grid_conf, final_dim, downsample):
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample, ogfW // downsample
    ds = torch.tensor(depth_bins(*grid_conf["ddiscr"], grid_conf["mode"]), dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = ds.shape[0]
    xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((xs, ys, ds), -1)


def geometry(fr, rots, trans, intrins, post_rots, post_trans):
    B, N, _ = trans.shape
    points = fr - post_trans.view(B, N, 1, 1, 1, 3)
    points = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(points.unsqueeze(-1))
    points = torch.cat((points[:, :, :, :, :, :2] * points[:, :, :, :, :, 2:3], points[:, :, :, :, :, 2:3]), 5)
    combine = rots.matmul(torch.inverse(intrins))
    points = combine.view(B, N, 1, 1, 1, 3, 3).matmul(points).squeeze(-1)
    return points + trans.view(B, N, 1, 1, 1, 3)


@torch.no_grad()
def lift_splat(depth_logit, x_img, rots, trans, intrins, post_rots, post_trans, grid_conf, final_dim, downsample):
    """depth_logit (B*N, D, fH, fW), x_img (B*N, C, fH, fW) -> BEV (B, C*nz, ny, nx): soft-max over depth, outer product with
    the features, and the sum of all frustum points that fall into each voxel.  The sum is an exact index_add in float64 (the
    reference's sort + float32 cumsum + difference computes the same sums with cancellation noise)."""
    B, N, _ = trans.shape
    BN, C, fH, fW = x_img.shape
    dx, bx, nx = gen_dx_bx(grid_conf["xbound"], grid_conf["ybound"], grid_conf["zbound"])
    fr = frustum(grid_conf, final_dim, downsample)
    D = fr.shape[0]
    geom = geometry(fr, rots, trans, intrins, post_rots, post_trans)                       # (B,N,D,fH,fW,3)
    depth = torch.softmax(depth_logit, dim=1)
    x = (depth.unsqueeze(1) * x_img.unsqueeze(2)).view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)   # (B,N,D,fH,fW,C)
    idx = ((geom - (bx - dx / 2.)) / dx).long().view(-1, 3)                                 # truncation toward zero, like .long()
    bix = torch.arange(B).view(B, 1).expand(B, N * D * fH * fW).reshape(-1)
    kept = ((idx[:, 0] >= 0) & (idx[:, 0] < nx[0]) & (idx[:, 1] >= 0) & (idx[:, 1] < nx[1]) & (idx[:, 2] >= 0) & (idx[:, 2] < nx[2]))
    flat = ((bix * nx[2] + idx[:, 2]) * nx[1] + idx[:, 1]) * nx[0] + idx[:, 0]             # (b, z, y, x)
    out = torch.zeros(B * int(nx[2]) * int(nx[1]) * int(nx[0]), C, dtype=torch.float64)
    out.index_add_(0, flat[kept], x.reshape(-1, C)[kept].double())
    out = out.view(B, int(nx[2]), int(nx[1]), int(nx[0]), C).permute(0, 4, 1, 2, 3)        # (B, C, nz, ny, nx)
    return torch.cat(out.unbind(dim=2), 1).float()                                          # collapse z like the reference
