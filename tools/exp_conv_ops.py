"""Per-launch timing of every distinct conv launch of one inference step (graph of 10 back-to-back launches each, so CPU
launch latency stays out of the number).  Usage: python tools/exp_conv_ops.py [scenes_per_step]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coalign_b200 import synth  # noqa: E402
from coalign_b200.engine import CoAlignEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12
args = synth.opv2v_args()
sd = synth.random_state_dict(args, 0)
rl = [5] * B
CM = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = CoAlignEngine(args, sd, sum(rl), len(rl), precise=False, use_graph=False, block_n_cap=256, chan_major_256=CM)
print("chan_major_256 =", CM)
scenes = [synth.make_scene(s, 5, 60000, args["lidar_range"], pose_noise=True) for s in range(B)]
pts = torch.from_numpy(np.concatenate([p for sc in scenes for p in sc["points"]])).cuda()
off = np.arange(0, sum(rl) + 1, dtype=np.int32) * 60000
pw = torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()
eng.forward_points(pts, off, rl, pw)
torch.cuda.synchronize()
ops = eng.build_descs(sum(rl), len(rl))
sp = torch.cuda.current_stream().cuda_stream
seen = {}
total = 0.0
for kind, o in ops:
    if kind != "conv":
        continue
    key = (o.n_img, o.Hp, o.Wp, o.n_total, o.n_ksteps * 64, o.block_n, o.out_mode, bool(o.residual), eng._has_tap_triples(o))
    if key not in seen:
        eng._launch_ops([(kind, o)], B, sp)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            cs = torch.cuda.current_stream().cuda_stream
            for _ in range(10):
                eng._launch_ops([(kind, o)], B, cs)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 30 * 1e3
        flops = 2.0 * o.n_img * (o.Hp - 2 * o.in_pad if hasattr(o, "in_pad") and o.in_pad else o.Hp - 2) * (o.Wp - 2) * o.n_total * o.n_ksteps * 64
        seen[key] = [us, flops, 0]
    seen[key][2] += 1
    total += seen[key][0]
print(f"{'imgs':>4} {'Hp':>4} {'Wp':>4} {'N':>4} {'K':>5} {'bn':>4} {'out':>3} {'res':>3} {'3x3':>3} {'count':>5} {'us':>8} {'TFLOP/s':>8}")
for k, (us, fl, cnt) in seen.items():
    print(f"{k[0]:4d} {k[1]:4d} {k[2]:4d} {k[3]:4d} {k[4]:5d} {k[5]:4d} {k[6]:3d} {int(k[7]):3d} {int(k[8]):3d} {cnt:5d} {us:8.1f} {fl / us / 1e6:8.0f}")
print(f"sum over the step's conv launches: {total / 1e3:.3f} ms")
