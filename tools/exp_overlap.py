"""Feasibility experiment (development aid): does the pillar front-end of the NEXT step overlap with the backbone of the
current one when they run on two streams?  Timing only (the concurrent front-end clobbers the canvas)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coalign_b200 import synth
from coalign_b200.engine import CoAlignEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
args = synth.opv2v_args(); sd = synth.random_state_dict(args, 0); rl = [5] * B
eng = CoAlignEngine(args, sd, sum(rl), len(rl), precise=False, block_n_cap=256)
scenes = [synth.make_scene(s, 5, 60000, args["lidar_range"], pose_noise=True) for s in range(B)]
pts = torch.from_numpy(np.concatenate([p for sc in scenes for p in sc["points"]])).cuda()
off = np.arange(0, sum(rl) + 1, dtype=np.int32) * 60000
pw = torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()
N = 50
def timeit(fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N
t_full = timeit(lambda: eng.forward_points(pts, off, rl, pw, clone=False))
eng._set_scene_meta(tuple(rl), pw)
t_bb = timeit(lambda: eng._graphed(("bb", tuple(rl)), tuple(rl)))
t_front = timeit(lambda: eng.run_front_only(off))
# front-end captured as its own graph on a side stream
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    eng.run_front_only(off)
    side.synchronize()
    with torch.cuda.graph(g, stream=side):
        eng.run_front_only(off)
main = torch.cuda.current_stream()
def both():
    side.wait_stream(main)
    with torch.cuda.stream(side):
        g.replay()
    eng._graphed(("bb", tuple(rl)), tuple(rl))
    main.wait_stream(side)
t_both = timeit(both)
print(f"B={B}: full step {t_full:.3f} ms | backbone only {t_bb:.3f} | front only {t_front:.3f} | backbone || front {t_both:.3f} ms "
      f"-> {B / t_both * 1e3:.0f} scenes/s (serial {B / t_full * 1e3:.0f})")
