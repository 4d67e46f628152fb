"""Synthetic replay of the reference's loop bodies on the B200 path (SURVEY 8c "can train.py / inference.py literally run
here?": no - datasets, open3d, spconv are absent - so the drop-in is demonstrated on synthetic `batch['ego']` dicts):

  inference.py:123-143 + inference_utils.inference_intermediate_fusion:   output = model(batch['ego']);
                                                                           boxes, scores = post_processor.post_process(...)
  train.py:105-125 (the whole iteration, on the device):                 model.train(); model.zero_grad(); optimizer.zero_grad()
                                                                           output = model(batch['ego'])
                                                                           loss = criterion(output, label_dict)
                                                                           loss.backward(); optimizer.step()

With the reference tree on PYTHONPATH (build container) the model and the loss are created by the reference's own
registries from a reference yaml with only `core_method` changed; without it (GPU box) the same classes are instantiated
directly.  Everything on the device runs in libcoalign_b200.so.

    python tools/demo_loop.py [--scenes 2] [--agents 3] [--points 20000]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import coalign_b200                                            # noqa: E402
from coalign_b200 import synth                                 # noqa: E402
from coalign_b200.loss import PointPillarLossB200              # noqa: E402
from coalign_b200.postprocess import VoxelPostprocessorB200    # noqa: E402


def build(use_registry: bool):
    args = synth.opv2v_args()
    if use_registry:
        from opencood.tools import train_utils                 # the unmodified reference
        coalign_b200.register()
        hypes = {"model": {"core_method": "point_pillar_coalign_b200", "args": args},
                 "loss": {"core_method": "point_pillar_loss_b200", "args": synth.loss_args()}}
        return args, train_utils.create_model(hypes), train_utils.create_loss(hypes)
    return args, coalign_b200.PointPillarCoalignB200(args), PointPillarLossB200(synth.loss_args())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", type=int, default=2)
    ap.add_argument("--agents", type=int, default=3)
    ap.add_argument("--points", type=int, default=20000)
    opt = ap.parse_args()
    try:
        import opencood  # noqa: F401
        use_registry = True
    except ImportError:
        use_registry = False
    args, model, criterion = build(use_registry)
    model.load_state_dict(synth.random_state_dict(args, 0), strict=True)
    model = model.cuda().eval()
    post = VoxelPostprocessorB200(synth.post_params(), train=False)
    anchors = torch.from_numpy(post.generate_anchor_box())
    print(f"model via {'reference registry' if use_registry else 'direct instantiation'}: {type(model).__name__}, "
          f"loss {type(criterion).__name__}, post-processor {type(post).__name__}")
    for it in range(opt.scenes):
        scene = synth.make_scene(100 + it, opt.agents, opt.points, args["lidar_range"], pose_noise=True)
        pts = torch.from_numpy(np.concatenate(scene["points"]).astype(np.float32)).cuda()
        off = (np.arange(opt.agents + 1) * opt.points).astype(np.int32)
        pw = torch.from_numpy(scene["pairwise_t_matrix"][None]).cuda()
        t0 = time.perf_counter()
        with torch.no_grad():
            out = model.forward_points(pts, off, [opt.agents], pw)            # raw clouds in (GPU voxelisation)
            boxes, scores = post.post_process({"ego": {"transformation_matrix": torch.eye(4).cuda(), "anchor_box": anchors}},
                                              {"ego": out})
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        n_box = 0 if boxes is None else int(boxes.shape[0])
        # a label dict with generate_label's structure (random positives: this is a plumbing demo, not a detector)
        case = synth.loss_case(seed=it, n=1, H=out["cls_preds"].shape[2], W=out["cls_preds"].shape[3], n_pos=20)
        label = {"pos_equal_one": torch.from_numpy(case["pos"]), "neg_equal_one": torch.from_numpy(case["neg"]),
                 "targets": torch.from_numpy(case["tgt"])}
        heads = {k: v.detach().clone().requires_grad_(True) for k, v in out.items()}
        loss = criterion(heads, label)
        criterion.logging(0, it, opt.scenes)
        loss.backward()
        gn = float(sum((v.grad ** 2).sum() for v in heads.values()) ** 0.5)
        print(f"scene {it}: cls {tuple(out['cls_preds'].shape)}, {n_box} boxes after NMS, forward + post-process {dt:.1f} ms "
              f"(first call includes graph capture), |dL/dheads| = {gn:.4f}")
    # ---- train.py:105-125: a few iterations of the reference's loop body with torch.optim.Adam (train_utils.py:196-206)
    model.train()
    optimizer = torch.optim.Adam(model.parameters(), lr=2e-3, eps=1e-10, weight_decay=1e-4)
    scene = synth.make_scene(500, opt.agents, opt.points, args["lidar_range"], pose_noise=True)
    eng = model.engine(opt.agents, 1)
    off = (np.arange(opt.agents + 1) * opt.points).astype(np.int32)
    vf, vc, vn, _ = eng.voxelize(torch.from_numpy(np.concatenate(scene["points"]).astype(np.float32)).cuda(), off, 32, 32000)
    batch = {"processed_lidar": {"voxel_features": vf, "voxel_coords": vc, "voxel_num_points": vn},
             "record_len": torch.tensor([opt.agents]), "pairwise_t_matrix": torch.from_numpy(scene["pairwise_t_matrix"][None]).cuda()}
    H, W = out["cls_preds"].shape[2:]
    case = synth.loss_case(seed=9, n=1, H=H, W=W, n_pos=20)
    label = {"pos_equal_one": torch.from_numpy(case["pos"]).cuda(), "neg_equal_one": torch.from_numpy(case["neg"]).cuda(),
             "targets": torch.from_numpy(case["tgt"]).cuda()}
    for it in range(4):
        model.zero_grad()
        optimizer.zero_grad()
        output = model(batch)
        loss = criterion(output, label)
        criterion.logging(0, it, 4)
        loss.backward()
        optimizer.step()
    n_grad = sum(p.grad is not None for p in model.parameters())
    print(f"train loop: 4 iterations, {n_grad} parameters received gradients from the device backward")
    print("demo ok")


if __name__ == "__main__":
    main()
