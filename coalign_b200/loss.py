"""Host-side mirror of the reference's ``PointPillarLoss`` (first piece of the training step): same
constructor argument (the yaml ``loss.args`` block), same ``forward(output_dict, target_dict, suffix)`` contract and
``loss_dict`` as /root/reference/opencood/loss/point_pillar_loss.py:14-116, but the focal / smooth-L1 / direction losses
AND their gradients w.r.t. the head outputs are produced by one pass of ``libcoalign_b200.so`` (``cb_pointpillar_loss``)
instead of ~40 torch kernels + autograd.  No CPU fallback.

``forward`` returns the total loss as a 0-dim CUDA tensor; when the predictions require grad it is attached to the autograd
graph through a custom Function whose backward hands out the precomputed gradients, so ``loss.backward()`` works on any
module that produced the head outputs.  ``last_grads`` keeps d(total)/d(cls, reg, dir) of the last call.

Not covered: the optional ``iou`` branch (pcdet CUDA op, not used by the CoAlign yaml).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, total, g_cls, g_reg, g_dir, cls_preds, reg_preds, dir_preds):
        ctx.save_for_backward(g_cls, g_reg, g_dir if g_dir is not None else torch.empty(0, device=total.device))
        ctx.has_dir = g_dir is not None
        return total.clone()

    @staticmethod
    def backward(ctx, grad_out):
        g_cls, g_reg, g_dir = ctx.saved_tensors
        return (None, None, None, None, grad_out * g_cls, grad_out * g_reg, grad_out * g_dir if ctx.has_dir else None)


class PointPillarLossB200(torch.nn.Module):
    def __init__(self, args: dict):
        super().__init__()
        self.pos_cls_weight = args["pos_cls_weight"]
        self.cls = args["cls"]
        self.reg = args["reg"]
        self.dir = args.get("dir")
        if "iou" in args:
            raise NotImplementedError("the iou branch (pcdet op) is not part of the CoAlign loss")
        self.loss_dict: Dict[str, float] = {}
        self.last_grads: Dict[str, Optional[torch.Tensor]] = {}
        self._ws: Optional[torch.Tensor] = None

    @staticmethod
    def _label(t: torch.Tensor, dev) -> torch.Tensor:
        t = t.to(dev)
        if t.dtype not in (torch.float32, torch.float64):
            t = t.float()
        return t.contiguous()

    def forward(self, output_dict: dict, target_dict: dict, suffix: str = "") -> torch.Tensor:
        lib = _lib.load(check_device=True)
        od = output_dict
        cls_preds = od[f"psm{suffix}"] if f"psm{suffix}" in od else od[f"cls_preds{suffix}"]
        reg_preds = od[f"rm{suffix}"] if f"rm{suffix}" in od else od[f"reg_preds{suffix}"]
        dir_preds = None
        if self.dir:
            dir_preds = od[f"dm{suffix}"] if f"dm{suffix}" in od else od[f"dir_preds{suffix}"]
        dev = cls_preds.device
        if dev.type != "cuda":
            raise RuntimeError("coalign_b200 loss needs CUDA tensors (no CPU fallback)")
        n, A, H, W = cls_preds.shape                      # batch_size = pos_equal_one.shape[0] (point_pillar_loss.py:43-48)
        pos = self._label(target_dict["pos_equal_one"], dev)
        neg = self._label(target_dict["neg_equal_one"], dev)
        tgt = self._label(target_dict["targets"], dev)
        if not (pos.dtype == neg.dtype == tgt.dtype):
            pos, neg, tgt = pos.double(), neg.double(), tgt.double()
        if pos.numel() != n * H * W * A or tgt.numel() != n * H * W * A * 7 or reg_preds.shape != (n, 7 * A, H, W):
            raise ValueError("label / prediction shapes do not match")
        num_bins, dir_offset, yaw = 0, 0.0, None
        if dir_preds is not None:
            num_bins = int(self.dir["args"]["num_bins"])
            dir_offset = float(self.dir["args"]["dir_offset"])
            yaw = np.deg2rad(np.asarray(self.dir["args"]["anchor_yaw"], dtype=np.float64))     # :148
            if yaw.shape[0] != A or dir_preds.shape != (n, num_bins * A, H, W):
                raise ValueError("dir_preds / anchor_yaw do not match the anchor configuration")
        c = cls_preds.detach().float().contiguous()
        r = reg_preds.detach().float().contiguous()
        d = dir_preds.detach().float().contiguous() if dir_preds is not None else None
        need = int(lib.cb_pointpillar_loss_workspace_bytes(n, H, W, A))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = torch.empty(4, dtype=torch.float32, device=dev)
        g_cls, g_reg = torch.empty_like(c), torch.empty_like(r)
        g_dir = torch.empty_like(d) if d is not None else None
        _lib.check(lib.cb_pointpillar_loss(
            c.data_ptr(), r.data_ptr(), d.data_ptr() if d is not None else None, pos.data_ptr(), neg.data_ptr(),
            tgt.data_ptr(), 1 if pos.dtype == torch.float64 else 0, n, H, W, A, num_bins,
            float(self.pos_cls_weight), float(self.cls["alpha"]), float(self.cls["gamma"]), float(self.cls["weight"]),
            float(self.reg["sigma"]), float(self.reg["weight"]), float(self.dir["weight"]) if self.dir else 0.0, dir_offset,
            yaw.ctypes.data if yaw is not None else None, out.data_ptr(), g_cls.data_ptr(), g_reg.data_ptr(),
            g_dir.data_ptr() if g_dir is not None else None, self._ws.data_ptr(), self._ws.numel(),
            torch.cuda.current_stream(dev).cuda_stream), "cb_pointpillar_loss")
        self.last_grads = {"cls_preds": g_cls, "reg_preds": g_reg, "dir_preds": g_dir}
        vals = out.tolist()                               # the reference syncs here too (.item() per term, :95,111-113)
        self.loss_dict.update({"total_loss": vals[0], "reg_loss": vals[1], "cls_loss": vals[2]})
        if self.dir:
            self.loss_dict["dir_loss"] = vals[3]
        total = out[0]
        needs_grad = cls_preds.requires_grad or reg_preds.requires_grad or (dir_preds is not None and dir_preds.requires_grad)
        if needs_grad and torch.is_grad_enabled():
            return _LossFn.apply(total, g_cls, g_reg, g_dir, cls_preds, reg_preds, dir_preds)
        return total.clone()

    def logging(self, epoch, batch_id, batch_len, writer=None, suffix=""):
        """Same console line and tensorboard scalars as PointPillarLoss.logging (point_pillar_loss.py:169-199), called by
        train.py once per iteration."""
        total_loss = self.loss_dict.get("total_loss", 0)
        reg_loss = self.loss_dict.get("reg_loss", 0)
        cls_loss = self.loss_dict.get("cls_loss", 0)
        dir_loss = self.loss_dict.get("dir_loss", 0)
        iou_loss = self.loss_dict.get("iou_loss", 0)
        print("[epoch %d][%d/%d]%s || Loss: %.4f || Conf Loss: %.4f"
              " || Loc Loss: %.4f || Dir Loss: %.4f || IoU Loss: %.4f" % (
                  epoch, batch_id + 1, batch_len, suffix, total_loss, cls_loss, reg_loss, dir_loss, iou_loss))
        if writer is not None:
            step = epoch * batch_len + batch_id
            writer.add_scalar("Regression_loss" + suffix, reg_loss, step)
            writer.add_scalar("Confidence_loss" + suffix, cls_loss, step)
            writer.add_scalar("Dir_loss" + suffix, dir_loss, step)
            writer.add_scalar("Iou_loss" + suffix, iou_loss, step)
