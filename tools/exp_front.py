"""Development aid: time the pillar front-end alone (canvas clear + voxelise + PFN + scatter) for B scenes of 5 agents,
eager launches and as a CUDA graph (CB_NO_PDL=1 switches the programmatic dependent launches off)."""
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
eng.forward_points(pts, off, rl, pw, clone=False)          # fills pts_buf, workspace
torch.cuda.synchronize()
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t_eager = timeit(lambda: eng.run_front_only(off))
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    eng.run_front_only(off); side.synchronize()
    with torch.cuda.graph(g, stream=side):
        eng.run_front_only(off)
torch.cuda.synchronize()
t_graph = timeit(g.replay)
by = B * 5 * (60000 * 16 + 200 * 704 * 64 * 2)
print(f"front-end B={B}: eager {t_eager:.1f} us, graph {t_graph:.1f} us "
      f"-> {by / t_graph / 1e3:.0f} GB/s algorithmic")
