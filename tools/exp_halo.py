"""Halo-kernel experiment (development aid): for every (CB_HALO, CB_HALO_BO) variant, in a fresh process,
  1. small-config parity of the tensor-core path against the SIMT evaluation of the same descriptors (rel-L2 per stage),
  2. time of each distinct conv launch of a B-scene step,
  3. whole graph-replayed forward (scenes/s).
A variant whose pipeline breaks traps inside the kernel (bounded mbarrier spin) and only kills its own child process."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from coalign_b200 import synth
    from coalign_b200.engine import CoAlignEngine
    from tests import golden_cases as G
    from tests.test_parity_gpu import cuda_batch, engine_stages, rel_l2

    halo = os.environ.get("CB_HALO", "1") != "0"
    # 1. parity vs SIMT on the small config
    args = G.small_args("att")
    sd = synth.random_state_dict(args, 4)
    inp = G.small_case_inputs([3, 2], seed0=104)
    res = {}
    for simt in (True, False):
        eng = CoAlignEngine(args, sd, 5, 2, precise=False, simt_conv=simt, use_graph=False)
        eng.halo = halo and not simt
        out = eng.forward_voxels(*cuda_batch(inp))
        torch.cuda.synchronize()
        st = engine_stages(eng, 5, 2)
        st.update({k: v.cpu().numpy() for k, v in out.items()})
        res[simt] = st
        del eng
    print("parity vs SIMT:", " ".join(f"{k}={rel_l2(res[False][k], res[True][k]):.1e}" for k in res[True]), flush=True)
    # 2. per-conv timing, full size
    B = int(os.environ.get("CB_B", "6"))
    args = synth.opv2v_args()
    sd = synth.random_state_dict(args, 0)
    rl = [5] * B
    eng = CoAlignEngine(args, sd, sum(rl), len(rl), precise=False, block_n_cap=256, use_graph=False)
    eng.halo = halo
    scenes = [synth.make_scene(s, 5, 60000, args["lidar_range"], pose_noise=True) for s in range(B)]
    pts = torch.from_numpy(np.concatenate([p for sc in scenes for p in sc["points"]])).cuda()
    off = np.arange(0, sum(rl) + 1, dtype=np.int32) * 60000
    pw = torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()
    eng.forward_points(pts, off, rl, pw)
    torch.cuda.synchronize()
    ops = eng.build_descs(sum(rl), len(rl))
    sp = torch.cuda.current_stream().cuda_stream
    seen = {}
    for kind, o in ops:
        if kind != "conv":
            continue
        key = (o.n_img * (o.Hp - 2) * (o.Wp - 2), o.n_total, o.n_ksteps * 64, o.block_n, o.out_mode, bool(o.residual),
               CoAlignEngine._has_tap_triples(o))
        if key in seen:
            continue
        for _ in range(2):
            eng._launch_ops([(kind, o)], B, sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            eng._launch_ops([(kind, o)], B, sp)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 5 * 1e3
        flop = 2.0 * key[0] * key[1] * key[2]
        seen[key] = (t, flop / t / 1e6)
    print("conv us    :", " ".join(f"{v[0]:7.1f}" for v in seen.values()), flush=True)
    print("conv TFLOPs:", " ".join(f"{v[1]:7.0f}" for v in seen.values()), flush=True)
    if os.environ.get("CB_KEYS"):
        print("keys:", list(seen.keys()))
    # 3. whole forward, graph
    eng.use_graph = True
    for _ in range(3):
        eng.forward_points(pts, off, rl, pw, clone=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.forward_points(pts, off, rl, pw, clone=False)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 20
    print(f"forward: {t:.3f} ms/step -> {B / t * 1e3:.1f} scenes/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for name, envs in (("staged STG epilogue", {"CB_EPI_DIRECT": "0", "CB_KEYS": "1"}),
                           ("direct 256-bit store epilogue (default)", {})):
            env = dict(os.environ)
            env.update(envs)
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True,
                                     text=True, timeout=240)
                print(f"=== {name}: rc={out.returncode}\n{out.stdout.strip()}\n{out.stderr[-600:] if out.returncode else ''}",
                      flush=True)
            except subprocess.TimeoutExpired:
                print(f"=== {name}: TIMEOUT", flush=True)
