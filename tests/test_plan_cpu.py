"""CPU check of the host logic: the engine's launch plan (packed weights, PF/PS layouts, K-step tables, epilogue
modes), executed by the reference interpreter in tests/plan_interpreter.py, reproduces the reference's golden
outputs."""
import os

import numpy as np
import pytest
import torch

from coalign_b200 import synth
from tests import golden_cases as G
from tests import plan_interpreter as PI

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,fusion", [("model_small_att", "att"), ("model_small_max", "max")])
def test_launch_plan_reproduces_reference_golden(name, fusion):
    from coalign_b200.engine import CoAlignEngine
    g = np.load(os.path.join(GOLD, name + ".npz"))
    seed = int(g["seed"])
    args = G.small_args(fusion)
    sd = synth.random_state_dict(args, seed)
    rl = [int(v) for v in g["record_len"]]
    inp = G.small_case_inputs(rl, seed0=100 + seed)
    eng = CoAlignEngine(args, sd, sum(rl), len(rl), device="cpu", precise=True, plan_only=True)
    out = PI.run_plan(eng, sd, args, G.to_torch_batch(inp))
    for i in range(3):
        got = PI.act_to_nchw(eng.lvl[i]["out"], eng.lvl[i]["out"].n_cap)[:sum(rl)].numpy()
        ref = g[f"feat{i}"]
        assert np.abs(got - ref).max() <= 1e-3 * np.sqrt((ref * ref).mean()) + 1e-3 * np.abs(ref).max(), f"feat{i}"
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        a, b = out[k].numpy().astype(np.float64), g[k].astype(np.float64)
        err = np.abs(a - b)
        assert (err <= 1e-3 * np.abs(b) + 1e-3 * np.sqrt((b * b).mean())).all(), (k, err.max())
    with pytest.raises(RuntimeError):
        eng._launch_ops([], 1, 0)


def test_single_agent_plan_reproduces_reference_golden():
    """BASELINE configs[0] (single-agent `point_pillar`, BaseBEVBackbone, no fusion): the plain-backbone launch plan
    (BN eps 1e-3 folding, conv stacks without residuals, PS->PF copies instead of fusion) against the golden vectors
    of the unmodified reference model."""
    from coalign_b200.engine import CoAlignEngine
    g = np.load(os.path.join(GOLD, "model_single_plain.npz"))
    seed, n = int(g["seed"]), int(g["n_frames"])
    args = synth.single_args(G.SMALL_RANGE, G.SMALL_VOXEL)
    sd = synth.random_state_dict(args, seed, backbone="plain")
    inp = G.single_case_inputs(n, seed0=100 + seed)
    eng = CoAlignEngine(args, sd, n, n, device="cpu", precise=True, plan_only=True, backbone="plain", fusion=False)
    out = PI.run_plan(eng, sd, args, G.to_torch_batch(inp))
    for i in range(3):
        got = PI.act_to_nchw(eng.lvl[i]["out"], eng.lvl[i]["out"].n_cap)[:n].numpy()
        ref = g[f"feat{i}"]
        assert np.abs(got - ref).max() <= 1e-3 * np.sqrt((ref * ref).mean()) + 1e-3 * np.abs(ref).max(), f"feat{i}"
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        a, b = out[k].numpy().astype(np.float64), g[k].astype(np.float64)
        err = np.abs(a - b)
        assert (err <= 1e-3 * np.abs(b) + 1e-3 * np.sqrt((b * b).mean())).all(), (k, err.max())


def test_stage1_uncertainty_plan_reproduces_reference_golden():
    """SURVEY 8f row 4 (`point_pillar_uncertainty`): heads straight on the 384-channel decoded map, four head segments
    (cls, reg, unc, dir) in one 32-column GEMM - launch plan against the golden vectors of the unmodified reference."""
    from coalign_b200.engine import CoAlignEngine
    g = np.load(os.path.join(GOLD, "model_single_uncertainty.npz"))
    seed, n = int(g["seed"]), int(g["n_frames"])
    args = synth.uncertainty_args(G.SMALL_RANGE, G.SMALL_VOXEL)
    sd = synth.random_state_dict(args, seed, backbone="plain")
    inp = G.single_case_inputs(n, seed0=100 + seed)
    eng = CoAlignEngine(args, sd, n, n, device="cpu", precise=True, plan_only=True, backbone="plain", fusion=False)
    assert eng.head_names == ["cls_preds", "reg_preds", "unc_preds", "dir_preds"] and eng.head_cn == [2, 14, 6, 4]
    out = PI.run_plan(eng, sd, args, G.to_torch_batch(inp))
    for k in ("cls_preds", "reg_preds", "unc_preds", "dir_preds"):
        a, b = out[k].numpy().astype(np.float64), g[k].astype(np.float64)
        err = np.abs(a - b)
        assert (err <= 1e-3 * np.abs(b) + 1e-3 * np.sqrt((b * b).mean())).all(), (k, err.max())


def test_engine_batch_validation_and_agent_offsets_on_cpu():
    """Host logic of the engine that needs no device: batch-shape validation (same errors a malformed collate output would
    hit in the reference as shape mismatches deep inside regroup / warp) and the agent-offset prefix sums the fusion
    kernel consumes (regroup, fusion_in_one.py:21-24)."""
    import pytest
    import torch
    from coalign_b200.engine import CoAlignEngine
    args = synth.make_args(G.SMALL_RANGE, [0.4, 0.4, 4])
    sd = synth.random_state_dict(args, 0)
    eng = CoAlignEngine(args, sd, 6, 3, device="cpu", precise=False, plan_only=True)
    pw = torch.eye(4, dtype=torch.float64).repeat(2, 5, 5, 1, 1)
    eng._set_scene_meta((2, 3), pw)
    assert eng.agent_off.tolist() == [0, 2, 5, 5]                       # padded with the total: empty trailing scenes
    assert torch.equal(eng.pairwise[:2], pw)
    with pytest.raises(ValueError):
        eng._set_scene_meta((2, 3, 2), torch.eye(4, dtype=torch.float64).repeat(3, 5, 5, 1, 1))    # 7 agents > capacity 6
    with pytest.raises(ValueError):
        eng._set_scene_meta((1, 1, 1, 1), torch.eye(4, dtype=torch.float64).repeat(4, 5, 5, 1, 1))  # 4 scenes > capacity 3
    with pytest.raises(ValueError):
        eng._set_scene_meta((0, 2), pw)                                                            # empty scene
    with pytest.raises(ValueError):
        eng._set_scene_meta((6,), torch.eye(4, dtype=torch.float64).repeat(1, 5, 5, 1, 1))          # 6 agents > max_cav 5
    with pytest.raises(ValueError):
        eng._set_scene_meta((2, 3), torch.eye(4, dtype=torch.float64).repeat(2, 4, 4, 1, 1))        # wrong L
    with pytest.raises(RuntimeError):
        eng._launch_ops([], 1, 0)                                                                  # plan-only: no launches
    single = CoAlignEngine(synth.single_args(G.SMALL_RANGE, G.SMALL_VOXEL),
                           synth.random_state_dict(synth.single_args(G.SMALL_RANGE, G.SMALL_VOXEL), 0, backbone="plain"),
                           2, 2, device="cpu", plan_only=True, backbone="plain", fusion=False)
    single._set_scene_meta((1, 1), None)
    with pytest.raises(ValueError):
        single._set_scene_meta((2,), None)                                                         # one agent per frame


def test_wgrad_and_dgrad_as_row_shifted_gemms_over_the_padded_flat_index():
    """DESIGN §8 plan for the training step, checked numerically: in the PF layout (zero halo) the weight gradient of a
    3x3/s1 convolution is nine plain GEMMs over the flattened padded pixel index with one constant row shift per tap, and
    the input gradient is the forward's shifted-GEMM form on dZ with the per-tap weights transposed and the shifts negated."""
    import torch
    from torch.nn.grad import conv2d_input, conv2d_weight
    g = torch.Generator().manual_seed(0)
    n, cin, cout, H, W = 2, 8, 16, 5, 7
    x = torch.randn(n, cin, H, W, generator=g, dtype=torch.float64)
    dz = torch.randn(n, cout, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, generator=g, dtype=torch.float64)
    Wp = W + 2

    def to_pf(t):                                                       # (n,C,H,W) -> [n*(H+2)*(W+2)][C], zero halo
        p = torch.zeros(t.shape[0], H + 2, Wp, t.shape[1], dtype=t.dtype)
        p[:, 1:H + 1, 1:W + 1, :] = t.permute(0, 2, 3, 1)
        return p.reshape(-1, t.shape[1])

    def shifted(a, s):                                                  # rows q -> a[q + s], zero outside (TMA OOB fill)
        out = torch.zeros_like(a)
        if s >= 0:
            out[:a.shape[0] - s] = a[s:]
        else:
            out[-s:] = a[:s]
        return out

    xp, dzp = to_pf(x), to_pf(dz)
    ref_dw = conv2d_weight(x, w.shape, dz, padding=1)
    ref_dx = conv2d_input(x.shape, w, dz, padding=1)
    dxp = torch.zeros_like(xp)
    for r in range(3):
        for s in range(3):
            shift = (r - 1) * Wp + (s - 1)                              # the forward's row shift of tap (r, s)
            dw_tap = dzp.t() @ shifted(xp, shift)                       # [cout][cin]: one MN-major GEMM, K = all rows
            assert torch.allclose(dw_tap, ref_dw[:, :, r, s], atol=1e-10), (r, s)
            dxp += shifted(dzp, -shift) @ w[:, :, r, s]                 # [rows][cin]
    dx = dxp.reshape(n, H + 2, Wp, cin)[:, 1:H + 1, 1:W + 1, :].permute(0, 3, 1, 2)
    assert torch.allclose(dx, ref_dx, atol=1e-10)
