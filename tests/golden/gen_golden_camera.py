"""Golden vectors of the camera model's BEV half (BASELINE configs[4]): the UNMODIFIED reference class
`BevEncodeMSFusion` (/root/reference/opencood/models/sub_modules/lss_submodule.py:357-417) on a small splat-output-shaped
input.  Run in the build container only:  python tests/golden/gen_golden_camera.py  ->  tests/golden/camera_bev_small.npz.
Import-only stubs: tests/golden/_stubs (+ efficientnet_pytorch, which lss_submodule imports at module scope for CamEncode)."""
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
if not hasattr(_f, "Where2commFusion"):                       # lss_submodule imports a name the reference no longer defines
    _f.Where2commFusion = object
from opencood.models.sub_modules.lss_submodule import BevEncodeMSFusion      # noqa: E402

from coalign_b200 import synth                                # noqa: E402


def main():
    for name, core, rl, seed in (("camera_bev_small", "att_ms", [3, 2], 11), ("camera_bev_small_max", "max_ms", [2, 2], 12)):
        torch.manual_seed(seed)
        m = BevEncodeMSFusion({"core_method": core, "args": {"in_channels": 128, "voxel_size": [0.4, 0.4, 20]}})
        sd = synth.random_camera_bev_state_dict(seed)
        assert set(sd) == set(m.state_dict()), set(sd) ^ set(m.state_dict())
        m.load_state_dict(sd, strict=True)
        m.eval()
        x, pw = synth.camera_bev_case(rl, seed, hw=48)
        with torch.no_grad():
            xs, xf = m(torch.from_numpy(x), torch.tensor(rl), torch.from_numpy(pw))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), seed=seed, record_len=np.asarray(rl),
                            x_single=xs.numpy(), x_fuse=xf.numpy())
        print(name, xs.shape, xf.shape, float(xs.abs().mean()), float(xf.abs().mean()))


if __name__ == "__main__":
    main()
