"""SURVEY 8f row 2 (training step), oracle side: `oracle.coalign_oracle.forward_train` (train-mode BatchNorm, autograd) +
`oracle.loss_oracle` reproduce the loss, the gradient of EVERY parameter and the running-statistic updates of the
unmodified reference in `.train()` mode (tests/golden/gen_golden_train.py).  This pins the oracle the backward kernels
will be tested against; no CUDA code is involved yet."""
import os

import numpy as np
import torch

from coalign_b200 import synth
from oracle import coalign_oracle as O
from oracle import loss_oracle as LO
from tests import golden_cases as G

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BN_MOMENTUM = {"pillar_vfe.pfn_layers.0.norm": 0.01, "backbone.resnet.layer0.0.bn1": 0.1,
               "backbone.resnet.layer2.0.downsample.1": 0.1, "backbone.deblocks.2.1": 0.01}


def test_training_step_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "train_small.npz"))
    seed, record_len = int(g["seed"]), [int(v) for v in g["record_len"]]
    args = synth.make_args(G.SMALL_RANGE, [0.4, 0.4, 4])
    sd0 = synth.random_state_dict(args, seed)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v.clone())
          for k, v in sd0.items()}
    inp = G.small_case_inputs(record_len, seed0=100 + seed)
    out, stats = O.forward_train(sd, args, G.to_torch_batch(inp))
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        ref = g["out_" + k]
        assert np.abs(out[k].detach().numpy() - ref).max() <= 2e-4 * np.sqrt((ref * ref).mean()) + 1e-5, k
    B, _, H, W = out["cls_preds"].shape
    case = synth.loss_case(seed=seed, n=B, H=H, W=W, n_pos=6)
    total, parts = LO.pointpillar_loss(synth.loss_args(), out["cls_preds"], out["reg_preds"], out["dir_preds"],
                                       torch.from_numpy(case["pos"]), torch.from_numpy(case["neg"]), torch.from_numpy(case["tgt"]))
    assert abs(float(total.detach()) - float(g["total_loss"])) <= 1e-4 * float(g["total_loss"])
    for k in ("reg_loss", "cls_loss", "dir_loss"):
        assert abs(float(parts[k].detach()) - float(g[k])) <= 1e-4 * abs(float(g[k])) + 1e-7, k
    total.backward()
    # gradients of all 127 parameters: norm within 1e-3, sampled entries within 1e-3 of the tensor's rms gradient
    checked = 0
    for name, p in sd.items():
        if not (torch.is_tensor(p) and p.requires_grad):
            continue
        assert p.grad is not None, name
        gr = p.grad.double().flatten().numpy()
        ref_norm = float(g["gn_" + name])
        assert abs(np.sqrt((gr * gr).sum()) - ref_norm) <= 1e-3 * ref_norm + 1e-9, (name, np.sqrt((gr * gr).sum()), ref_norm)
        rms = ref_norm / np.sqrt(gr.size)
        idx = np.array([(gr.size * k) // 5 for k in (1, 2, 3, 4)])
        assert np.abs(gr[:4] - g["g4_" + name]).max() <= 1e-3 * rms + 1e-9, name
        assert np.abs(gr[idx] - g["gs_" + name]).max() <= 1e-3 * rms + 1e-9, name
        checked += 1
    assert checked == 127
    # running statistics after the step: (1 - m) * running + m * batch (unbiased variance), m per layer kind
    for pre, m in BN_MOMENTUM.items():
        mean, var = stats[pre]
        rm = (1 - m) * sd0[pre + ".running_mean"] + m * mean
        rv = (1 - m) * sd0[pre + ".running_var"] + m * var
        np.testing.assert_allclose(rm.numpy(), g["rm_" + pre], rtol=1e-4, atol=1e-6, err_msg=pre)
        np.testing.assert_allclose(rv.numpy(), g["rv_" + pre], rtol=1e-4, atol=1e-6, err_msg=pre)


def _loss_grads_fn(seed, dtype):
    def fn(out):                                        # the piece that already runs on the device (cb_pointpillar_loss)
        B, _, H, W = out["cls_preds"].shape
        case = synth.loss_case(seed=seed, n=B, H=H, W=W, n_pos=6)
        leaves = [out[k].detach().clone().requires_grad_(True) for k in ("cls_preds", "reg_preds", "dir_preds")]
        with torch.enable_grad():
            total, _ = LO.pointpillar_loss(synth.loss_args(), *leaves, torch.from_numpy(case["pos"]),
                                           torch.from_numpy(case["neg"]), torch.from_numpy(case["tgt"]))
            total.backward()
        return {k: t.grad.to(dtype) for k, t in zip(("cls_preds", "reg_preds", "dir_preds"), leaves)}
    return fn


def test_hand_written_backward_matches_autograd_and_reference_golden():
    """The training step written out by hand (oracle/backward_oracle.py: explicit dgrad / wgrad, train-mode BatchNorm
    reductions, soft-max / warp adjoints, PFN arg-max routing - the pieces the device kernels have to provide, no
    autograd): (1) in float64 it equals autograd of the pinned train-mode oracle to 1e-9 for all 127 parameters (the
    formulas are exact); (2) in float32 it reproduces the unmodified reference's gradients up to float32 round-off, which
    the small-batch BatchNorm of this tiny case (3 x 7 top-level maps) amplifies to a few 1e-3 of a tensor's rms gradient."""
    from oracle import backward_oracle as BO
    g = np.load(os.path.join(GOLD, "train_small.npz"))
    seed, record_len = int(g["seed"]), [int(v) for v in g["record_len"]]
    args = synth.make_args(G.SMALL_RANGE, [0.4, 0.4, 4])
    sd32 = synth.random_state_dict(args, seed)
    inp = G.small_case_inputs(record_len, seed0=100 + seed)
    batch = G.to_torch_batch(inp)
    # (1) float64: hand-written backward == autograd of forward_train
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd32.items()}
    leaves = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running_" not in k else v) for k, v in sd64.items()}
    out_a, _ = O.forward_train(leaves, args, batch)
    lg = _loss_grads_fn(seed, torch.float64)(out_a)
    torch.autograd.backward([out_a[k] for k in lg], [lg[k] for k in lg])
    with torch.no_grad():
        out_m, grads64 = BO.forward_backward(sd64, args, batch, _loss_grads_fn(seed, torch.float64))
    assert len(grads64) == 127
    for name, gm in grads64.items():
        ga = leaves[name].grad
        assert float((gm - ga).abs().max()) <= 1e-9 * float(ga.abs().max()) + 1e-14, name
    # (2) float32 against the reference's golden gradients
    with torch.no_grad():
        out, grads = BO.forward_backward(sd32, args, batch, _loss_grads_fn(seed, torch.float32))
    for k in ("cls_preds", "reg_preds", "dir_preds"):
        ref = g["out_" + k]
        assert np.abs(out[k].numpy() - ref).max() <= 2e-4 * np.sqrt((ref * ref).mean()) + 1e-5, k
    names = [k[3:] for k in g.files if k.startswith("gn_")]
    assert len(names) == 127 and set(names) == set(grads), sorted(set(names) ^ set(grads))[:5]
    for name in names:
        gr = grads[name].double().flatten().numpy()
        ref_norm = float(g["gn_" + name])
        assert abs(np.sqrt((gr * gr).sum()) - ref_norm) <= 3e-3 * ref_norm + 1e-9, (name, np.sqrt((gr * gr).sum()), ref_norm)
        rms = ref_norm / np.sqrt(gr.size)
        idx = np.array([(gr.size * k) // 5 for k in (1, 2, 3, 4)])
        assert np.abs(gr[:4] - g["g4_" + name]).max() <= 2e-2 * rms + 1e-9, name
        assert np.abs(gr[idx] - g["gs_" + name]).max() <= 2e-2 * rms + 1e-9, name
