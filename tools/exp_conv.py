"""Conv-kernel experiments: time each conv launch of a B=4 step under CB_DEBUG variants (development aid)."""
import os, sys, subprocess
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from coalign_b200 import synth
    from coalign_b200.engine import CoAlignEngine
    B = int(os.environ.get("CB_B", "4"))
    args = synth.opv2v_args(); sd = synth.random_state_dict(args, 0); rl = [5] * B
    eng = CoAlignEngine(args, sd, sum(rl), len(rl), precise=False, block_n_cap=256, use_graph=False, pair=os.environ.get("CB_PAIR", "0") == "1")
    scenes = [synth.make_scene(s, 5, 60000, args["lidar_range"], pose_noise=True) for s in range(B)]
    pts = torch.from_numpy(np.concatenate([p for sc in scenes for p in sc["points"]])).cuda()
    off = np.arange(0, sum(rl) + 1, dtype=np.int32) * 60000
    pw = torch.from_numpy(np.stack([sc["pairwise_t_matrix"] for sc in scenes])).cuda()
    eng.forward_points(pts, off, rl, pw); torch.cuda.synchronize()
    ops = eng.build_descs(sum(rl), len(rl)); sp = torch.cuda.current_stream().cuda_stream
    seen = {}
    for kind, o in ops:
        if kind != "conv": continue
        key = (o.n_img * (o.Hp - 2) * (o.Wp - 2), o.n_total, o.n_ksteps * 64, o.block_n, o.out_mode, bool(o.residual))
        if key in seen: continue
        for _ in range(2): eng._launch_ops([(kind, o)], B, sp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): eng._launch_ops([(kind, o)], B, sp)
        e1.record(); torch.cuda.synchronize()
        seen[key] = e0.elapsed_time(e1) / 5 * 1e3
    print(" ".join(f"{v:7.1f}" for v in seen.values()))
    if os.environ.get("CB_DEBUG", "0") == "0":
        print("keys:", list(seen.keys()))
else:
    for name, envs in (("pair base", {"CB_PAIR": "1"}), ("pair 3 stages", {"CB_PAIR": "1", "CB_DEBUG": "32"}),
                       ("pair L2 prefetch", {"CB_PAIR": "1", "CB_DEBUG": "16"}), ("pair no-epilogue", {"CB_PAIR": "1", "CB_DEBUG": "1"})):
        env = dict(os.environ, CB_DEBUG="0", CB_B="4")
        env.update(envs)
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(f"{name:18s}", out.stdout.strip().split("\n")[0], out.stderr[-300:] if out.returncode else "", flush=True)
