"""One training iteration of the CoAlign model on the device: the loop body of the reference's trainer
(/root/reference/opencood/tools/train.py:105-125 - model.train(); forward; criterion; backward; optimizer.step) with

  forward + backward   coalign_b200.train_engine.TrainEngine (train-mode BatchNorm, all kernels ours)
  criterion            cb_pointpillar_loss (loss + d loss / d heads in one pass; loss/point_pillar_loss.py:36-116)
  DDP all-reduce       the gradients live in ONE flat fp32 buffer ordered by backward completion; each bucket (heads +
                       shrink + deblocks | level 2 | level 1 | level 0 | PFN) is all-reduced with NCCL as soon as the backward
                       segment that produces it has been enqueued, overlapping the rest of the backward pass
                       (train_ddp.py:104-109: DistributedDataParallel's bucketed overlap; 12.9 M fp32 gradients = 51.6 MB)
  optimizer            cb_adam_step = torch.optim.Adam(lr, eps, weight_decay) of train_utils.setup_optimizer (:196-206)

`use_graph`: the forward(+loss) and every backward segment are captured into CUDA graphs once per batch signature; the NCCL
calls sit between the replays on the same stream order DDP uses (collectives wait for the producing segment, the next
segment does not wait for the collective).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .train_engine import TrainEngine


def allreduce_buckets(flat: torch.Tensor, buckets: Sequence, upto: Optional[int] = None, start: int = 0, async_op: bool = True):
    """All-reduce (sum) the slices buckets[start:upto] of the flat gradient buffer, in order; returns the work handles.
    Bucket i holds the gradients the i-th backward segment completes, so calling this right after enqueueing segment i
    overlaps the collective with segments i+1.. (the role of DistributedDataParallel's reducer, train_ddp.py:104-109)."""
    import torch.distributed as dist
    works = []
    for a, b in buckets[start:upto]:
        w = dist.all_reduce(flat[a:b], async_op=async_op)
        if async_op:
            works.append(w)
    return works


class Trainer:
    def __init__(self, args: dict, state_dict: Dict[str, torch.Tensor], loss_args: dict, max_agents: int, max_scenes: int,
                 max_voxels_total: int = 0, lr: float = 2e-3, eps: float = 1e-10, weight_decay: float = 1e-4,
                 betas=(0.9, 0.999), precise: bool = False, device="cuda", use_graph: bool = True, max_cav: int = 5,
                 distributed: bool = False):
        self.eng = TrainEngine(args, state_dict, max_agents, max_scenes, device=device, precise=precise, max_cav=max_cav,
                               max_voxels_total=max_voxels_total)
        self.lib = self.eng.lib
        self.loss_args = loss_args
        if "iou" in loss_args:
            raise NotImplementedError("the iou branch (pcdet op) is not part of the CoAlign loss")
        self.lr, self.eps, self.wd, self.betas = float(lr), float(eps), float(weight_decay), betas
        self.use_graph = use_graph
        self.distributed = bool(distributed)
        self.world = 1
        if self.distributed:
            import torch.distributed as dist
            self.world = dist.get_world_size()
        dev = self.eng.device
        n = self.eng.n_flat
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss_out = torch.zeros(4, dtype=torch.float32, device=dev)
        self.A = int(args["anchor_number"])
        H0, W0, _ = self.eng.levels[0]
        self.H0, self.W0 = H0, W0
        NS = self.eng.max_scenes
        self._labels: Optional[dict] = None
        need = int(self.lib.cb_pointpillar_loss_workspace_bytes(NS, H0, W0, self.A))
        self._loss_ws = torch.empty(need, dtype=torch.uint8, device=dev)
        d = loss_args.get("dir")
        self._yaw = np.deg2rad(np.asarray(d["args"]["anchor_yaw"], dtype=np.float64)) if d else None
        self._graphs: Dict[tuple, dict] = {}
        self.comm_ms = 0.0
        self.last_overlap = None

    # ------------------------------------------------------------------ pieces
    def _label_bufs(self, labels: dict):
        dt = torch.float64 if labels["pos_equal_one"].dtype == torch.float64 else torch.float32
        if self._labels is None or self._labels["dtype"] != dt:
            NS, H, W, A = self.eng.max_scenes, self.H0, self.W0, self.A
            dev = self.eng.device
            self._labels = {"dtype": dt, "pos": torch.zeros(NS, H, W, A, dtype=dt, device=dev),
                            "neg": torch.zeros(NS, H, W, A, dtype=dt, device=dev),
                            "tgt": torch.zeros(NS, H, W, 7 * A, dtype=dt, device=dev)}
            self._graphs.clear()
        return self._labels

    def _loss(self, n_sc: int):
        e, la = self.eng, self.loss_args
        lb = self._labels
        names = e.head_names
        cls_o, reg_o = e.head_out[names.index("cls_preds")], e.head_out[names.index("reg_preds")]
        cls_g, reg_g = e.head_grad[names.index("cls_preds")], e.head_grad[names.index("reg_preds")]
        has_dir = "dir_preds" in names and la.get("dir")
        dir_o = e.head_out[names.index("dir_preds")] if has_dir else None
        dir_g = e.head_grad[names.index("dir_preds")] if has_dir else None
        d = la.get("dir") or {}
        _lib.check(self.lib.cb_pointpillar_loss(
            cls_o.data_ptr(), reg_o.data_ptr(), dir_o.data_ptr() if has_dir else None, lb["pos"].data_ptr(),
            lb["neg"].data_ptr(), lb["tgt"].data_ptr(), 1 if lb["dtype"] == torch.float64 else 0, n_sc, self.H0, self.W0, self.A,
            int(d["args"]["num_bins"]) if has_dir else 0, float(la["pos_cls_weight"]), float(la["cls"]["alpha"]),
            float(la["cls"]["gamma"]), float(la["cls"]["weight"]), float(la["reg"]["sigma"]), float(la["reg"]["weight"]),
            float(d["weight"]) if has_dir else 0.0, float(d["args"]["dir_offset"]) if has_dir else 0.0,
            self._yaw.ctypes.data if has_dir else None, self.loss_out.data_ptr(), cls_g.data_ptr(), reg_g.data_ptr(),
            dir_g.data_ptr() if has_dir else None, self._loss_ws.data_ptr(), self._loss_ws.numel(),
            torch.cuda.current_stream(e.device).cuda_stream), "cb_pointpillar_loss")

    def _adam(self):
        e = self.eng
        _lib.check(self.lib.cb_adam_step(e.pflat.data_ptr(), e.gflat.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), e.n_flat,
                                         self.lr, self.betas[0], self.betas[1], self.eps, self.wd, 1.0 / self.world,
                                         self.step_dev.data_ptr(), 1, torch.cuda.current_stream(e.device).cuda_stream),
                   "cb_adam_step")

    def _segments(self, ent, n_sc: int):
        """[callable] : segment 0 = forward + loss + backward up to the first bucket mark, then one per bucket."""
        segs: List[list] = [[]]
        for op in ent["bwd"]:
            segs[-1].append(op)
            if op[0] == "bucket":
                segs.append([])
        if not segs[-1]:
            segs.pop()
        fns = []
        for i, ops in enumerate(segs):
            if i == 0:
                def f0(ops=ops):
                    self.eng.run_ops(ent["fwd"])
                    self._loss(n_sc)
                    self.eng.run_ops(ops)
                fns.append(f0)
            else:
                fns.append(lambda ops=ops: self.eng.run_ops(ops))
        return fns

    # ------------------------------------------------------------------ one iteration
    def step(self, batch: dict, labels: dict) -> torch.Tensor:
        """batch: reference-schema tensors on the device (voxel_features, voxel_coords, voxel_num_points, record_len,
        pairwise_t_matrix); labels: pos_equal_one / neg_equal_one / targets of label_dict.  Returns the total loss (0-dim
        device tensor, no host sync)."""
        e = self.eng
        rl = e.set_batch(batch["voxel_features"], batch["voxel_coords"], batch["voxel_num_points"], batch["record_len"],
                         batch["pairwise_t_matrix"])
        n_sc = len(rl)
        lb = self._label_bufs(labels)
        for k, src in (("pos", "pos_equal_one"), ("neg", "neg_equal_one"), ("tgt", "targets")):
            lb[k][:n_sc].copy_(labels[src].to(lb["dtype"]).reshape(lb[k][:n_sc].shape), non_blocking=True)
        g = self._graphs.get(rl)
        if g is None:
            g = {"fns": self._segments(e.plan(rl), n_sc), "graphs": None, "runs": 0}
            self._graphs[rl] = g
        if self.use_graph and g["graphs"] is None and g["runs"] >= 1:
            torch.cuda.synchronize(e.device)
            cap = torch.cuda.Stream(device=e.device)
            graphs = []
            for fn in g["fns"]:
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg, stream=cap):
                    fn()
                graphs.append(cg)
            g["graphs"] = graphs
        works = []
        for i in range(len(g["fns"])):
            if g["graphs"] is not None:
                g["graphs"][i].replay()
            else:
                g["fns"][i]()
            if self.distributed and i < len(e.buckets):
                works += allreduce_buckets(e.gflat, e.buckets, upto=i + 1, start=i)
        for w in works:
            w.wait()                                   # stream-level wait: the optimizer runs after the collectives
        self._adam()
        g["runs"] += 1
        e.num_batches_tracked += 1
        return self.loss_out[0]

    def set_lr(self, lr: float):
        self.lr = float(lr)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return self.eng.state_dict_out()
