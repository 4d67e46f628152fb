"""Pin the CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from coalign_b200 import synth
from oracle import coalign_oracle as O
from tests import golden_cases as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _close(a, b, rtol=1e-4, atol=1e-4):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert (err <= tol).all(), f"max err {err.max():.3e}, worst ratio {(err / tol).max():.2f}"


def test_pose_and_affine_ops():
    g = np.load(os.path.join(GOLD, "ops.npz"))
    pw = O.pairwise_from_poses(list(g["poses"]), 5)
    np.testing.assert_allclose(pw, g["pairwise"], rtol=0, atol=1e-12)
    aff = O.normalize_pairwise_tfm(torch.from_numpy(g["pairwise"][None]), 200, 704, 0.4)
    np.testing.assert_allclose(aff.numpy(), g["affine_200_704"], rtol=0, atol=1e-13)
    # the product-side synthetic generator builds the same matrices
    sc = synth.make_scene(3, 4, 10, [-140.8, -40, -3, 140.8, 40, 1])
    np.testing.assert_allclose(sc["pairwise_t_matrix"], O.pairwise_from_poses(sc["poses"], 5), atol=1e-12)


def test_warp_golden():
    g = np.load(os.path.join(GOLD, "ops.npz"))
    out = O.warp_affine_simple(torch.from_numpy(g["warp_src"]), torch.from_numpy(g["warp_M"]))
    _close(out.numpy(), g["warp_out"], rtol=1e-5, atol=2e-6)


def test_fusion_golden():
    g = np.load(os.path.join(GOLD, "ops.npz"))
    x, aff = torch.from_numpy(g["att_x"]), torch.from_numpy(g["att_affine"])
    _close(O.att_fusion(x, [3, 2], aff, "att").numpy(), g["att_out"], rtol=1e-5, atol=2e-6)
    _close(O.att_fusion(x, [3, 2], aff, "max").numpy(), g["max_out"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name,fusion", [("model_small_att", "att"), ("model_small_single", "att"),
                                         ("model_small_max", "max")])
def test_model_golden(name, fusion):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed = int(g["seed"])
    args = G.small_args(fusion)
    sd = synth.random_state_dict(args, seed)
    inp = G.small_case_inputs([int(v) for v in g["record_len"]], seed0=100 + seed)
    # the (unpinned) voxel stage must at least reproduce what the fixture was generated from
    assert np.array_equal(inp["voxel_coords"], g["voxel_coords"])
    assert np.array_equal(inp["voxel_num_points"], g["voxel_num_points"])
    assert abs(inp["voxel_features"].astype(np.float64).sum() - float(g["voxel_features_sum"])) < 1e-6
    stages = {}
    out = O.forward(sd, args, G.to_torch_batch(inp), stages)
    _close(stages["pillar_features"].numpy(), g["pillar_features"], rtol=1e-4, atol=1e-4)
    for i in range(3):
        _close(stages["feats"][i].numpy(), g[f"feat{i}"], rtol=1e-3, atol=1e-3)
        _close(stages["fused"][i].numpy(), g[f"fused{i}"], rtol=1e-3, atol=1e-3)
    _close(stages["shrunk"].numpy(), g["shrunk"], rtol=1e-3, atol=1e-3)
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        _close(out[k].numpy(), g[k], rtol=1e-3, atol=1e-3)


def test_voxelizer_c_vs_python_loop():
    from oracle import voxelize_np as V
    rng = np.random.default_rng(5)
    pts = synth.lidar_cloud(rng, 1500, G.SMALL_RANGE, sigma=4.0)
    # points exactly on cell edges / range limits (boundary KAT)
    edge = np.array([[-11.2, -4.8, -3.0, 0.5], [11.2, 0, 0, 0.5], [0, 4.8, 0, 0.5], [0, 0, 1.0, 0.5],
                     [0.4, 0.8, -1.0, 0.1], [11.199999, 4.799999, 0.999, 0.2], [-11.2000001, 0, 0, 0.3]], np.float32)
    pts = np.concatenate([edge, pts])
    for max_pts, max_vox in ((32, 70000), (3, 70000), (32, 100)):
        a = V.voxelize_c(pts, G.SMALL_RANGE, G.SMALL_VOXEL, max_pts, max_vox)
        b = V.voxelize_py(pts, G.SMALL_RANGE, G.SMALL_VOXEL, max_pts, max_vox)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert a[0].shape[0] <= max_vox and a[2].max() <= max_pts
    empty = V.voxelize_c(np.zeros((0, 4), np.float32), G.SMALL_RANGE, G.SMALL_VOXEL, 32, 10)
    assert empty[0].shape == (0, 32, 4)


def test_single_agent_pointpillar_golden():
    """BASELINE configs[0]: the single-agent `point_pillar` model (BaseBEVBackbone, no fusion) - oracle restatement
    against the unmodified reference created through its yaml + registry (tests/golden/gen_golden_single.py)."""
    g = np.load(os.path.join(GOLD, "model_single_plain.npz"))
    seed, n = int(g["seed"]), int(g["n_frames"])
    args = synth.single_args(G.SMALL_RANGE, G.SMALL_VOXEL)
    sd = synth.random_state_dict(args, seed, backbone="plain")
    inp = G.single_case_inputs(n, seed0=100 + seed)
    assert np.array_equal(inp["voxel_coords"], g["voxel_coords"])
    assert np.array_equal(inp["voxel_num_points"], g["voxel_num_points"])
    st = {}
    out = O.forward_single(sd, args, G.to_torch_batch(inp), st)
    for i in range(3):
        _close(st["feats"][i].numpy(), g[f"feat{i}"])
    _close(st["decoded"].numpy(), g["decoded"])
    _close(st["shrunk"].numpy(), g["shrunk"])
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        _close(out[k].numpy(), g[k])


def test_stage1_uncertainty_detector_golden():
    """the stage-1 `point_pillar_uncertainty` detector (BaseBEVBackbone, no shrink header, extra
    unc_head) - oracle restatement against the unmodified reference created through its yaml + registry
    (tests/golden/gen_golden_single.py::main_uncertainty)."""
    g = np.load(os.path.join(GOLD, "model_single_uncertainty.npz"))
    seed, n = int(g["seed"]), int(g["n_frames"])
    args = synth.uncertainty_args(G.SMALL_RANGE, G.SMALL_VOXEL)
    sd = synth.random_state_dict(args, seed, backbone="plain")
    assert sd["unc_head.weight"].shape == (6, 384, 1, 1) and "shrink_conv.layers.0.double_conv.0.weight" not in sd
    inp = G.single_case_inputs(n, seed0=100 + seed)
    st = {}
    out = O.forward_single(sd, args, G.to_torch_batch(inp), st)
    _close(st["decoded"].numpy(), g["decoded"])
    assert list(out) == ["cls_preds", "reg_preds", "unc_preds", "dir_preds"]      # point_pillar_uncertainty.py:68-76
    for k in out:
        _close(out[k].numpy(), g[k])
