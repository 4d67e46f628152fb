"""GPU diagnostics (development aid, lives under tests/ because it uses the oracle): per-stage error tables and
per-op timings.  Run under gpurun: python tests/dev_diag_gpu.py [small] [full]."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from coalign_b200 import synth, _lib          # noqa: E402
from coalign_b200.engine import CoAlignEngine  # noqa: E402
from tests import golden_cases as G           # noqa: E402
from tests.test_parity_gpu import cuda_batch, engine_stages, oracle_stages, rel_l2   # noqa: E402


def stats(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    e = np.abs(a - b)
    return f"rel_l2={rel_l2(a, b):.2e} max_abs={e.max():.2e} rms_ref={np.sqrt((b*b).mean()):.2e} argmax={np.unravel_index(e.argmax(), e.shape)}"


def small_case():
    seed = 1
    args = G.small_args("att")
    sd = synth.random_state_dict(args, seed)
    rl = [3, 2]
    inp = G.small_case_inputs(rl, seed0=100 + seed)
    ref_out, ref_st = oracle_stages(args, sd, inp)
    ref = {"canvas": ref_st["canvas"].numpy(), "decoded": ref_st["decoded"].numpy(), "shrunk": ref_st["shrunk"].numpy()}
    for i in range(3):
        ref[f"feat{i}"] = ref_st["feats"][i].numpy()
        ref[f"fused{i}"] = ref_st["fused"][i].numpy()
    for k, v in ref_out.items():
        ref[k] = v.numpy()
    res = {}
    for precise in (True, False):
        for simt in (True, False):
            eng = CoAlignEngine(args, sd, 5, 2, precise=precise, simt_conv=simt, use_graph=False)
            out = eng.forward_voxels(*cuda_batch(inp))
            torch.cuda.synchronize()
            st = engine_stages(eng, 5, 2)
            st.update({k: v.cpu().numpy() for k, v in out.items()})
            res[(precise, simt)] = st
            print(f"--- precise={precise} simt={simt} vs fp32 oracle")
            for k in st:
                print(f"  {k:10s} {stats(st[k], ref[k])}")
    for precise in (True, False):
        print(f"--- precise={precise}: TC vs SIMT")
        for k in res[(precise, True)]:
            print(f"  {k:10s} {stats(res[(precise, False)][k], res[(precise, True)][k])}")


def full_case(n_agents=5, n_scenes=1, precise=False, block_n=128, n_points=60000, check=False, pair=True):
    args = synth.opv2v_args()
    sd = synth.random_state_dict(args, 0)
    rl = [n_agents] * n_scenes
    eng = CoAlignEngine(args, sd, sum(rl), len(rl), precise=precise, block_n_cap=block_n, use_graph=False, pair=pair)
    scenes = [synth.make_scene(s, n_agents, n_points, args["lidar_range"], pose_noise=True) for s in range(n_scenes)]
    pts = np.concatenate([p for sc in scenes for p in sc["points"]])
    off = np.arange(0, sum(rl) + 1, dtype=np.int32) * n_points
    pw = torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()
    pts_d = torch.from_numpy(pts).cuda()
    eng.forward_points(pts_d, off, rl, pw)
    torch.cuda.synchronize()
    # per-op timing (eager)
    ops = eng.build_descs(sum(rl), len(rl))
    sp = torch.cuda.current_stream().cuda_stream
    import ctypes as C
    lib = eng.lib
    tot_flop, tot_t = 0.0, 0.0
    print(f"=== full size: agents={n_agents} scenes={n_scenes} precise={precise} block_n_cap={block_n} pair={pair}")
    for kind, o in ops:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 5
        for _ in range(2):
            eng._launch_ops([(kind, o)], len(rl), sp)
        ev[0].record()
        for _ in range(reps):
            eng._launch_ops([(kind, o)], len(rl), sp)
        ev[1].record()
        torch.cuda.synchronize()
        t = ev[0].elapsed_time(ev[1]) / reps * 1e-3
        if kind == "conv":
            rows = o.n_img * (o.Hp - 2) * (o.Wp - 2)
            flop = 2.0 * rows * o.n_total * o.n_ksteps * 64 / (3 if precise else 1)
            tot_flop += flop; tot_t += t
            print(f"  conv rows={rows:7d} N={o.n_total:4d} K={o.n_ksteps*64:5d} bn={o.block_n:3d} mode={o.out_mode} "
                  f"t={t*1e6:8.1f}us  {flop/t/1e12:7.1f} TFLOP/s(alg)")
        else:
            h, w, c = eng.levels[o]
            byts = (n_agents + 1) * n_scenes * h * w * c * 2
            print(f"  fuse level {o}: t={t*1e6:8.1f}us  {byts/t/1e9:7.1f} GB/s(alg)")
            tot_t += t
    print(f"  conv total {tot_flop/1e9:.1f} GFLOP in {tot_t*1e3:.3f} ms -> {tot_flop/tot_t/1e12:.1f} TFLOP/s; "
          f"scenes/s (backbone only) = {n_scenes/tot_t:.1f}")
    # whole forward, graph
    eng.use_graph = True
    for _ in range(3):
        eng.forward_points(pts_d, off, rl, pw, clone=False)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(10):
        eng.forward_points(pts_d, off, rl, pw, clone=False)
    ev[1].record()
    torch.cuda.synchronize()
    t = ev[0].elapsed_time(ev[1]) / 10 * 1e-3
    print(f"  whole forward (graph): {t*1e3:.3f} ms/step -> {n_scenes/t:.1f} scenes/s")
    if check:
        from oracle import coalign_oracle as O
        from tests.golden_cases import scenes_to_batch, to_torch_batch
        torch.set_num_threads(os.cpu_count())
        t0 = time.time()
        inp = scenes_to_batch(scenes, args["lidar_range"], args["voxel_size"])
        t1 = time.time()
        ref = O.forward(sd, args, to_torch_batch(inp))
        t2 = time.time()
        print(f"  oracle: voxelize {t1-t0:.2f}s forward {t2-t1:.2f}s on {os.cpu_count()} threads")
        out = eng.forward_points(pts_d, off, rl, pw)
        for k in ref:
            print(f"  {k}: {stats(out[k].cpu().numpy(), ref[k].numpy())}")


if __name__ == "__main__":
    what = sys.argv[1:] or ["small", "full"]
    if "small" in what:
        small_case()
    if "full" in what:
        full_case(5, 4, precise=False, block_n=256, pair=False)
        full_case(5, 4, precise=False, block_n=256, pair=True, check=False)
        full_case(5, 4, precise=False, block_n=128, pair=True)
        full_case(5, 1, precise=False, block_n=256, pair=True)
        full_case(5, 8, precise=False, block_n=256, pair=True)
        full_case(2, 1, precise=True, block_n=256, pair=True, check=True)
