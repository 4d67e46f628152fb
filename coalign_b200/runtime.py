"""Host <-> device pipelining around the engine: the public end-to-end API for serving raw point clouds.

`PipelinedRunner` overlaps, across consecutive steps, (1) the host->device copy of the next batch's points and
poses from pinned memory, (2) the forward of the current batch (one CUDA graph) and (3) the device->host copy of
the previous batch's cls/reg/dir maps, using a copy-in stream, the compute stream and a copy-out stream with
double-buffered staging on both sides.  Every step still moves all of its inputs and outputs over PCIe; nothing is
cached between steps.

The reference does the same job serially (`train_utils.to_device` then `model(batch)` then `.cpu()` in
post-processing; /root/reference/opencood/tools/inference.py:125-143, utils/box_utils.py:714-715).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .engine import CoAlignEngine


class PipelinedRunner:
    def __init__(self, engine: CoAlignEngine, max_points: int, depth: int = 2):
        self.eng = engine
        dev = engine.device
        self.depth = depth
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.s_cmp = torch.cuda.Stream(device=dev)
        B, L = engine.max_scenes, engine.max_cav
        self.d_pts = [torch.empty(max_points, 4, dtype=torch.float32, device=dev) for _ in range(depth)]
        self.d_pw = [torch.empty(B, L, L, 4, 4, dtype=torch.float64, device=dev) for _ in range(depth)]
        self.d_out = [[torch.empty_like(t) for t in engine.head_out] for _ in range(depth)]
        self.h_out = [[torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in engine.head_out] for _ in range(depth)]
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_cmp = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]
        self.ev_free_in = [torch.cuda.Event() for _ in range(depth)]     # compute finished reading input slot
        self.ev_free_out = [torch.cuda.Event() for _ in range(depth)]    # copy-out finished reading d_out slot
        self._n = 0
        self._meta: List[Optional[int]] = [None] * depth

    def submit(self, host_points: torch.Tensor, pt_offset: Sequence[int], record_len: Sequence[int],
               host_pairwise: torch.Tensor, max_pts: int = 32, max_voxels: int = 70000) -> int:
        """Enqueue one batch (pinned host tensors).  Returns a ticket for `result`."""
        k = self._n % self.depth
        n_scenes = len(record_len)
        total = int(pt_offset[-1])
        eng = self.eng
        if self._n >= self.depth:
            self.s_in.wait_event(self.ev_free_in[k])             # slot's previous contents consumed
            self.s_cmp.wait_event(self.ev_free_out[k])           # slot's previous outputs copied out
        with torch.cuda.stream(self.s_in):
            self.d_pts[k][:total].copy_(host_points[:total], non_blocking=True)
            self.d_pw[k][:n_scenes].copy_(host_pairwise[:n_scenes], non_blocking=True)
            self.ev_in[k].record(self.s_in)
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(self.ev_in[k])
            eng.forward_points(self.d_pts[k], pt_offset, record_len, self.d_pw[k][:n_scenes], max_pts, max_voxels,
                               clone=False)
            self.ev_free_in[k].record(self.s_cmp)
            for dst, src in zip(self.d_out[k], eng.head_out):
                dst[:n_scenes].copy_(src[:n_scenes], non_blocking=True)
            self.ev_cmp[k].record(self.s_cmp)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_cmp[k])
            for dst, src in zip(self.h_out[k], self.d_out[k]):
                dst[:n_scenes].copy_(src[:n_scenes], non_blocking=True)
            self.ev_out[k].record(self.s_out)
            self.ev_free_out[k].record(self.s_out)
        self._meta[k] = n_scenes
        t = self._n
        self._n += 1
        return t

    def result(self, ticket: int) -> Dict[str, torch.Tensor]:
        """Block until the ticket's outputs are in pinned host memory; valid until `depth` more submits."""
        k = ticket % self.depth
        self.ev_out[k].synchronize()
        n = self._meta[k]
        return {name: t[:n] for name, t in zip(self.eng.head_names, self.h_out[k])}

    def drain(self):
        self.s_out.synchronize()
        self.s_cmp.synchronize()
