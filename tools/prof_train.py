"""One full-size training iteration (4 scenes x 5 agents, OPV2V shape) eagerly, for ncu launch lists:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_train.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coalign_b200 import synth
from coalign_b200.engine import CoAlignEngine
from coalign_b200.trainer import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
args = synth.opv2v_args(); sd = synth.random_state_dict(args, 0)
NA, P = 5, 60000
eng = CoAlignEngine(args, sd, B * NA, B, precise=False, block_n_cap=256, use_graph=False)
scenes = [synth.make_scene(s, NA, P, args["lidar_range"], pose_noise=True) for s in range(B)]
pts = torch.from_numpy(np.concatenate([p for sc in scenes for p in sc["points"]])).cuda()
off = np.arange(0, B * NA + 1, dtype=np.int32) * P
vf, vc, vn, _ = eng.voxelize(pts, off, 32, 32000)
batch = {"voxel_features": vf, "voxel_coords": vc, "voxel_num_points": vn, "record_len": [NA] * B,
         "pairwise_t_matrix": torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()}
del eng
case = synth.loss_case(seed=0, n=B, H=100, W=352, n_pos=40)
labels = {"pos_equal_one": torch.from_numpy(case["pos"]).cuda(), "neg_equal_one": torch.from_numpy(case["neg"]).cuda(),
          "targets": torch.from_numpy(case["tgt"]).cuda()}
tr = Trainer(args, sd, synth.loss_args(), max_agents=B * NA, max_scenes=B, max_voxels_total=int(vf.shape[0]) + 1024, use_graph=False)
tr.step(batch, labels)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(batch, labels)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
