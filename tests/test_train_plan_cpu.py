"""CPU check of the training step's host logic: the launch plan of coalign_b200.train_engine.TrainEngine -
forward GEMM descriptors with raw weights, train-mode BatchNorm maps, input-gradient GEMMs (negated shifts, parity-plane
launches with the fused 1x1 downsample term), cb_wgrad unit / box lists, packed-weight and gradient permutations, buffer
wiring - executed by tests/train_plan_interpreter.py in exact fp32, reproduces the gradients of all 127 parameters, the
head outputs and the running-statistic updates of the UNMODIFIED reference in `.train()` mode (tests/golden/train_small.npz)
to the float32 noise floor of this network (two fp32 evaluations of the same graph differ by 2-5e-3 in the encoder
gradients: oracle fp32 vs fp64)."""
import os

import numpy as np
import pytest
import torch

from coalign_b200 import synth
from tests import golden_cases as G
from tests import train_plan_interpreter as TPI
from tests.test_train_oracle_cpu import BN_MOMENTUM, _loss_grads_fn

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _case():
    g = np.load(os.path.join(GOLD, "train_small.npz"))
    seed, record_len = int(g["seed"]), [int(v) for v in g["record_len"]]
    args = synth.make_args(G.SMALL_RANGE, [0.4, 0.4, 4])
    sd = synth.random_state_dict(args, seed)
    inp = G.small_case_inputs(record_len, seed0=100 + seed)
    return g, seed, record_len, args, sd, inp


def test_training_plan_reproduces_reference_gradients_and_running_stats():
    from coalign_b200.train_engine import TrainEngine
    g, seed, rl, args, sd, inp = _case()
    eng = TrainEngine(args, sd, sum(rl), len(rl), device="cpu", plan_only=True, exact_fp32_plan=True,
                      max_voxels_total=inp["voxel_features"].shape[0])
    with torch.no_grad():
        out, grads = TPI.run_train_plan(eng, args, G.to_torch_batch(inp), _loss_grads_fn(seed, torch.float32))
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        ref = g["out_" + k]
        assert np.abs(out[k].numpy() - ref).max() <= 2e-4 * np.sqrt((ref * ref).mean()) + 1e-5, k
    names = [k[3:] for k in g.files if k.startswith("gn_")]
    assert len(names) == 127 and set(names) == set(grads)
    for name in names:
        gr = grads[name].double().flatten().numpy()
        ref_norm = float(g["gn_" + name])
        assert abs(np.sqrt((gr * gr).sum()) - ref_norm) <= 1e-2 * ref_norm + 1e-9, (name, np.sqrt((gr * gr).sum()), ref_norm)
        rms = ref_norm / np.sqrt(gr.size)
        idx = np.array([(gr.size * k) // 5 for k in (1, 2, 3, 4)])
        assert np.abs(gr[:4] - g["g4_" + name]).max() <= 4e-2 * rms + 1e-9, name
        assert np.abs(gr[idx] - g["gs_" + name]).max() <= 4e-2 * rms + 1e-9, name
    for pre in BN_MOMENTUM:
        np.testing.assert_allclose(eng.R[pre + ".running_mean"].numpy(), g["rm_" + pre], rtol=1e-4, atol=1e-5, err_msg=pre)
        np.testing.assert_allclose(eng.R[pre + ".running_var"].numpy(), g["rv_" + pre], rtol=1e-4, atol=1e-5, err_msg=pre)
    # the flat gradient buffer is the concatenation of the parameters in backward-completion order (all-reduce buckets)
    assert eng.buckets[0][0] == 0 and eng.buckets[-1][1] == eng.n_flat
    assert all(a[1] == b[0] for a, b in zip(eng.buckets[:-1], eng.buckets[1:]))
    with pytest.raises(RuntimeError):
        eng.run_ops([])


def test_training_plan_max_fusion_matches_hand_written_backward():
    """fusion_method: max (MaxFusion, fusion_in_one.py:83-86) through the same plan, against oracle/backward_oracle.py."""
    from coalign_b200.train_engine import TrainEngine
    from oracle import backward_oracle as BO
    seed, rl = 5, [2, 3]
    args = G.small_args("max")
    sd = synth.random_state_dict(args, seed)
    inp = G.small_case_inputs(rl, seed0=400)
    batch = G.to_torch_batch(inp)
    with torch.no_grad():
        _out_ref, g_ref = BO.forward_backward(sd, args, batch, _loss_grads_fn(seed, torch.float32))
    eng = TrainEngine(args, sd, 5, 2, device="cpu", plan_only=True, exact_fp32_plan=True,
                      max_voxels_total=inp["voxel_features"].shape[0])
    with torch.no_grad():
        _out, grads = TPI.run_train_plan(eng, args, batch, _loss_grads_fn(seed, torch.float32))
    for name, b in g_ref.items():
        a = grads[name].double()
        b = b.double()
        rel = float((a - b).norm() / (b.norm() + 1e-30))
        assert rel < 3e-2, (name, rel)


def test_permute_job_strides_match_the_layout_semantics():
    """The (rows, K, strides) records handed to cb_permute_f32 / the batched cb_pack_job table reproduce the permutation the
    plan interpreter applies by layout (packed GEMM order -> the parameter's own layout), for both job kinds."""
    import torch
    from coalign_b200.train_engine import TrainEngine
    from tests.train_plan_interpreter import emu_permute
    rng = torch.Generator().manual_seed(0)
    for q in ({"kind": "conv", "cout": 6, "cin": 5, "taps": 9}, {"kind": "deconv", "cin": 7, "cu": 4, "k": 2},
              {"kind": "conv", "cout": 3, "cin": 8, "taps": 1}):
        n = (q["cout"] * q["cin"] * q["taps"]) if q["kind"] == "conv" else (q["cin"] * q["cu"] * q["k"] ** 2)
        src = torch.randn(n + 11, generator=rng)
        shape = (q["cout"], q["cin"], 3, 3) if q.get("taps") == 9 else (
            (q["cout"], q["cin"], 1, 1) if q["kind"] == "conv" else (q["cin"], q["cu"], q["k"], q["k"]))
        want = torch.zeros(shape)
        emu_permute(None, dict(q, src=src, src_off=11, dst=want))
        R1, R0, K0, (s_r1, s_r0, s_k1, s_k0) = TrainEngine.permute_job_shape(q)
        got = torch.empty(R1 * R0 * K0)
        for r1 in range(R1):
            for r0 in range(R0):
                for k0 in range(K0):
                    got[(r1 * R0 + r0) * K0 + k0] = src[11 + r1 * s_r1 + r0 * s_r0 + k0 * s_k0]
        assert torch.equal(got.view(shape), want), q
