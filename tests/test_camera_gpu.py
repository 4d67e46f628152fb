"""GPU parity of the camera model's BEV half: BevEncoderEngine through the C ABI against golden vectors of
the UNMODIFIED reference `BevEncodeMSFusion` (tests/golden/camera_bev_small*.npz) - precise mode within 1e-3, bf16 at bf16
drift - plus the two new layout kernels against PyTorch."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from coalign_b200 import _lib, synth
from tests.test_camera_cpu import CASES, _load

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("name,method", CASES)
def test_camera_bev_encoder_matches_reference_golden(name, method):
    from coalign_b200.camera import BevEncoderEngine
    g, rl, sd, x, pw = _load(name)
    for precise in (True, False):
        eng = BevEncoderEngine(sd, 48, 48, sum(rl), len(rl), discrete_ratio=0.4, method=method, precise=precise)
        for rep in range(2):                              # second call replays the captured graph
            xs, xf = eng.forward(x.cuda(), rl, pw.cuda())
        torch.cuda.synchronize()
        for got, key in ((xs, "x_single"), (xf, "x_fuse")):
            ref = g[key]
            got = got.cpu().numpy()
            if precise:
                err = np.abs(got - ref)
                assert (err <= 1e-3 * np.abs(ref) + 1e-3 * np.sqrt((ref * ref).mean())).all(), (key, err.max())
            else:
                assert rel_l2(got, ref) < 3e-2, (key, rel_l2(got, ref))


def test_upsample_concat_matches_torch():
    """cb_upsample_concat == nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) + torch.cat (lss_submodule.py:
    23-24,36-37) on bf16-rounded inputs."""
    lib = _lib.load(True)
    sp = torch.cuda.current_stream().cuda_stream
    torch.manual_seed(0)
    n, h, w, c1, c2 = 2, 6, 9, 64, 128
    a = torch.randn(n, c1, 2 * h, 2 * w, device="cuda").to(torch.bfloat16).float()
    b = torch.randn(n, c2, h, w, device="cuda").to(torch.bfloat16).float()

    def to_pf(t):
        nn_, cc, hh, ww = t.shape
        p = torch.zeros(nn_, hh + 2, ww + 2, cc, dtype=torch.bfloat16, device="cuda")
        p[:, 1:-1, 1:-1] = t.permute(0, 2, 3, 1).to(torch.bfloat16)
        return p.reshape(-1, cc).contiguous()
    pa, pb_ = to_pf(a), to_pf(b)
    dst = torch.zeros(n * (2 * h + 2) * (2 * w + 2), c1 + c2, dtype=torch.bfloat16, device="cuda")
    _lib.check(lib.cb_upsample_concat(pa.data_ptr(), 0, n, 2 * h, 2 * w, c1, 1, dst.data_ptr(), 0, c1 + c2, 0, sp))
    _lib.check(lib.cb_upsample_concat(pb_.data_ptr(), 0, n, h, w, c2, 2, dst.data_ptr(), 0, c1 + c2, c1, sp))
    torch.cuda.synchronize()
    got = dst.float().reshape(n, 2 * h + 2, 2 * w + 2, c1 + c2)[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
    ref = torch.cat([a, F.interpolate(b, scale_factor=2, mode="bilinear", align_corners=True)], 1)
    assert torch.equal(got[:, :c1], a)
    assert (got[:, c1:] - ref[:, c1:]).abs().max().item() < 2e-2          # one bf16 rounding of the interpolated value
    assert dst.reshape(n, 2 * h + 2, 2 * w + 2, -1)[:, 0].abs().max().item() == 0


def test_bev_encode_ms_fusion_twin_matches_engine_and_golden():
    """The nn.Module twin of the reference sub-module (same constructor argument, same 118 state_dict entries, same forward
    signature) in precise mode against the reference golden."""
    from coalign_b200.camera import BevEncodeMSFusionB200
    g, rl, sd, x, pw = _load("camera_bev_small")
    m = BevEncodeMSFusionB200({"core_method": "att_ms", "args": {"in_channels": 128, "voxel_size": [0.4, 0.4, 20],
                                                               "b200_precise": True}})
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        xs, xf = m(x.cuda(), torch.tensor(rl).cuda(), pw.cuda())
    torch.cuda.synchronize()
    for got, key in ((xs, "x_single"), (xf, "x_fuse")):
        ref = g[key]
        err = np.abs(got.cpu().numpy() - ref)
        assert (err <= 1e-3 * np.abs(ref) + 1e-3 * np.sqrt((ref * ref).mean())).all(), (key, err.max())
    with pytest.raises(NotImplementedError):
        m.train()(x.cuda(), torch.tensor(rl).cuda(), pw.cuda())


def test_lift_splat_matches_reference_golden_and_feeds_the_bev_encoder():
    """cb_lift_splat against the unmodified reference's lift + voxel_pooling output: frustum points whose float32 geometry
    lands within rounding of a voxel boundary may switch voxels between two evaluations of the same formulas, so the bound
    is on the map (rel-L2 < 2e-3, 99.9 % of the cells within 1e-3 of the map's scale), not bit-exactness; then the channels-
    last accumulator goes straight into the BEV encoder."""
    from coalign_b200.camera import BevEncoderEngine, LiftSplatB200
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lift_splat_small.npz"))
    case = synth.lift_splat_case(seed=int(g["seed"]))
    t = {k: torch.from_numpy(v).cuda() for k, v in case.items() if isinstance(v, np.ndarray)}
    ls = LiftSplatB200(case["grid_conf"], case["final_dim"], case["downsample"])
    acc = ls(t["depth_logit"], t["x_img"], t["rots"], t["trans"], t["intrins"], t["post_rots"], t["post_trans"])
    torch.cuda.synchronize()
    got = acc.permute(0, 3, 1, 2).cpu().numpy()
    ref = g["bev"]
    assert got.shape == ref.shape
    scale = np.abs(ref).max()
    assert rel_l2(got, ref) < 2e-3, rel_l2(got, ref)
    assert (np.abs(got - ref) <= 1e-3 * scale).mean() > 0.999
    # lift + splat -> BEV encoder without leaving the device (64 channels here: a 64-channel stem)
    sd = synth.random_camera_bev_state_dict(3, in_channels=64)
    eng = BevEncoderEngine(sd, 80, 80, 2, 1, discrete_ratio=0.4, method="att")
    pw = torch.eye(4, dtype=torch.float64).repeat(1, 5, 5, 1, 1).cuda()
    xs1, xf1 = eng.forward(acc, [2], pw, channels_last=True)
    xs2, xf2 = eng.forward(acc.permute(0, 3, 1, 2).contiguous(), [2], pw)
    torch.cuda.synchronize()
    assert torch.equal(xs1, xs2) and torch.equal(xf1, xf2) and torch.isfinite(xf1).all()
