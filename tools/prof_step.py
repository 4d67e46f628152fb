"""One full-size forward_points step, eager (no graph), for ncu captures of the non-conv kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coalign_b200 import synth
from coalign_b200.engine import CoAlignEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
args = synth.opv2v_args(); sd = synth.random_state_dict(args, 0); rl = [5] * B
eng = CoAlignEngine(args, sd, sum(rl), len(rl), precise=False, block_n_cap=256, use_graph=False)
scenes = [synth.make_scene(s, 5, 60000, args["lidar_range"], pose_noise=True) for s in range(B)]
pts = torch.from_numpy(np.concatenate([p for sc in scenes for p in sc["points"]])).cuda()
off = np.arange(0, sum(rl) + 1, dtype=np.int32) * 60000
pw = torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()
for _ in range(3):
    eng.forward_points(pts, off, rl, pw)
torch.cuda.synchronize()
