"""Golden vectors of lift + splat: the UNMODIFIED reference methods LiftSplatShoot.create_frustum / get_geometry /
voxel_pooling (/root/reference/opencood/models/lift_splat_shoot.py:64-169) called on a stand-in `self` that carries exactly
the attributes they read (the class's __init__ builds an EfficientNet and moves tensors to CUDA, neither available here), and
the lift lines of CamEncode.forward / get_cam_feats restated inline (soft-max over depth, outer product, view + permute).
Run in the build container only:  python tests/golden/gen_golden_lift.py  ->  tests/golden/lift_splat_small.npz."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference", ROOT]
eff = types.ModuleType("efficientnet_pytorch")
eff.EfficientNet = type("EfficientNet", (), {})
sys.modules["efficientnet_pytorch"] = eff

import numpy as np      # noqa: E402
import torch            # noqa: E402

import opencood.models.fuse_modules.fusion_in_one as _f      # noqa: E402
if not hasattr(_f, "Where2commFusion"):
    _f.Where2commFusion = object
from opencood.models.lift_splat_shoot import LiftSplatShoot          # noqa: E402
from opencood.utils.camera_utils import gen_dx_bx                    # noqa: E402

from coalign_b200 import synth                                        # noqa: E402


def main():
    case = synth.lift_splat_case(seed=21)
    gc, fd, ds_ = case["grid_conf"], case["final_dim"], case["downsample"]
    me = types.SimpleNamespace(data_aug_conf={"final_dim": fd}, downsample=ds_, grid_conf=gc, use_quickcumsum=True)
    me.dx, me.bx, me.nx = gen_dx_bx(gc["xbound"], gc["ybound"], gc["zbound"])
    me.frustum = LiftSplatShoot.create_frustum(me)
    t = {k: torch.from_numpy(v) for k, v in case.items() if isinstance(v, np.ndarray)}
    with torch.no_grad():
        geom = LiftSplatShoot.get_geometry(me, t["rots"], t["trans"], t["intrins"], t["post_rots"], t["post_trans"])
        depth = torch.softmax(t["depth_logit"], dim=1)                                   # CamEncode.get_depth_dist
        new_x = depth.unsqueeze(1) * t["x_img"].unsqueeze(2)                            # CamEncode.forward: the lift
        B, N = t["trans"].shape[:2]
        C, D = t["x_img"].shape[1], depth.shape[1]
        x = new_x.view(B, N, C, D, new_x.shape[-2], new_x.shape[-1]).permute(0, 1, 3, 4, 5, 2)   # get_cam_feats
        bev = LiftSplatShoot.voxel_pooling(me, geom, x)
    np.savez_compressed(os.path.join(HERE, "lift_splat_small.npz"), seed=21, bev=bev.numpy())
    print("bev", tuple(bev.shape), float(bev.abs().mean()), float((bev != 0).float().mean()))


if __name__ == "__main__":
    main()
