"""Development aid: time the three warp+attention fusion launches of one step (B scenes of 5 agents, OPV2V shape), eager
launches with CUDA events, inputs = the level outputs of a real forward.  Env: CB_FUSE_V (8 = the 8-channel-per-lane kernel), CB_FUSE_BLEND=32 (fp32 tap blend), CB_FUSE_OCC=3."""
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
eng.forward_points(pts, off, rl, pw, clone=False)
torch.cuda.synchronize()
ops = [(k, o) for k, o in eng.build_descs(sum(rl), B) if k == "fuse"]
sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = [0.0] * len(ops)
reps = 10
for r in range(reps + 2):
    for i, op in enumerate(ops):
        flush.zero_()                                    # maps come from HBM, as inside a step (> L2 of activations)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng._launch_ops([op], B, sp); b.record()
        torch.cuda.synchronize()
        if r >= 2: tot[i] += a.elapsed_time(b) * 1e3 / reps
by = 6 * 3942400 * 2 * B
print(f"fuse V={os.environ.get('CB_FUSE_V','9')} blend={os.environ.get('CB_FUSE_BLEND','16')} occ={os.environ.get('CB_FUSE_OCC','4')}: "
      + " + ".join(f"{t:.1f}" for t in tot) + f" = {sum(tot):.1f} us -> {by / sum(tot) / 1e3:.0f} GB/s algorithmic")
