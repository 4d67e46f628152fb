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
