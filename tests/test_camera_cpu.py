"""CPU tests of the camera model's BEV half (BASELINE configs[4]): (1) the oracle restatement of
`BevEncodeMSFusion` against golden vectors of the UNMODIFIED reference class; (2) the B200 engine's launch plan (7x7/s2 stem
over the pad-2 PS layout incl. its K-split in precise mode, BasicBlocks, up-sample + concat ops, decoder convs) executed by
the plan interpreter against the same golden vectors."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from coalign_b200 import synth
from oracle import camera_oracle as CO
from oracle import coalign_oracle as O
from tests import plan_interpreter as PI

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("camera_bev_small", "att"), ("camera_bev_small_max", "max")]


def _load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed, rl = int(g["seed"]), [int(v) for v in g["record_len"]]
    sd = synth.random_camera_bev_state_dict(seed)
    x, pw = synth.camera_bev_case(rl, seed, hw=48)
    return g, rl, sd, torch.from_numpy(x), torch.from_numpy(pw)


@pytest.mark.parametrize("name,method", CASES)
def test_camera_bev_oracle_matches_reference_golden(name, method):
    g, rl, sd, x, pw = _load(name)
    xs, xf = CO.bev_encode_ms_fusion(sd, x, rl, pw, 0.4, method)
    np.testing.assert_allclose(xs.numpy(), g["x_single"], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(xf.numpy(), g["x_fuse"], rtol=1e-4, atol=2e-4)


def run_bev_plan(eng, x, rl, pw, method):
    n_img, n_sc = sum(rl), len(rl)
    full = torch.zeros(eng.x_in.n_cap, eng.in_c, eng.ny, eng.nx)
    full[:n_img] = x
    PI.nchw_to_act(full, eng.x_in)
    affine = O.normalize_pairwise_tfm(pw, eng.ny, eng.nx, eng.voxel_size[0])
    for kind, o in eng.build_bev_ops(n_img, n_sc):
        if kind == "conv":
            PI.run_conv(eng, o)
        elif kind == "fuse":
            src, dst = eng.lvl[o]["out"], eng.lvl[o]["fused"]
            fused = O.att_fusion(PI.act_to_nchw(src, src.n_cap)[:n_img], rl, affine, method)
            fz = torch.zeros(dst.n_cap, dst.C, dst.H, dst.W)
            fz[:n_sc] = fused
            PI.nchw_to_act(fz, dst)
        elif kind == "pscopy":
            L = eng.lvl[o["li"]]
            PI.nchw_to_act(PI.act_to_nchw(L["out"], L["out"].n_cap), L["out_pf"])
        elif kind == "ups":
            s, d, n = o["src"], o["dst"], o["n"]
            v = PI.act_to_nchw(s, s.n_cap)[:n, :o["c"]]
            if o["scale"] == 2:
                v = F.interpolate(v, scale_factor=2, mode="bilinear", align_corners=True)
            nn_, hh, ww = torch.meshgrid(torch.arange(n), torch.arange(v.shape[2]), torch.arange(v.shape[3]), indexing="ij")
            PI._store(d, PI.act_rows(d, nn_, hh, ww).reshape(-1), o["ch"], v.permute(0, 2, 3, 1).reshape(-1, o["c"]))
        else:
            raise RuntimeError(kind)
    return PI.act_to_nchw(eng.dec["x_single"], n_img), PI.act_to_nchw(eng.dec["x_fuse"], n_sc)


@pytest.mark.parametrize("name,method", CASES)
def test_camera_bev_launch_plan_reproduces_reference_golden(name, method):
    from coalign_b200.camera import BevEncoderEngine
    g, rl, sd, x, pw = _load(name)
    eng = BevEncoderEngine(sd, 48, 48, sum(rl), len(rl), discrete_ratio=0.4, method=method, device="cpu", precise=True,
                           plan_only=True)
    xs, xf = run_bev_plan(eng, x, rl, pw, method)
    for got, key in ((xs, "x_single"), (xf, "x_fuse")):
        ref = g[key]
        err = np.abs(got.numpy() - ref)
        assert (err <= 1e-3 * np.abs(ref) + 1e-3 * np.sqrt((ref * ref).mean())).all(), (key, err.max())


def test_lift_splat_oracle_matches_reference_golden():
    """oracle.camera_oracle.lift_splat (exact float64 index_add) against the UNMODIFIED reference's create_frustum /
    get_geometry / voxel_pooling (tests/golden/gen_golden_lift.py); the reference's float32 cumsum trick carries its own
    cancellation noise, hence 1e-4 of the map's scale."""
    g = np.load(os.path.join(GOLD, "lift_splat_small.npz"))
    case = synth.lift_splat_case(seed=int(g["seed"]))
    t = {k: torch.from_numpy(v) for k, v in case.items() if isinstance(v, np.ndarray)}
    bev = CO.lift_splat(t["depth_logit"], t["x_img"], t["rots"], t["trans"], t["intrins"], t["post_rots"], t["post_trans"],
                        case["grid_conf"], case["final_dim"], case["downsample"])
    ref = g["bev"]
    assert bev.shape == ref.shape
    assert np.abs(bev.numpy() - ref).max() <= 1e-4 * np.abs(ref).max()
    assert (ref != 0).mean() > 0.05                                        # the case actually fills a part of the grid
