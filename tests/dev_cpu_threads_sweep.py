"""Thread-count sweep of the CPU arm (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from coalign_b200 import synth
args, batches = bench.make_batches(1, 1, 0)
sd = synth.random_state_dict(args, 0)
for th in (16, 32, 64, 128):
    bench.cpu_scene_seconds(args, sd, batches[0][2][0], th)
    t = [bench.cpu_scene_seconds(args, sd, batches[0][2][0], th)[0] for _ in range(2)]
    print(th, "threads:", t, flush=True)
