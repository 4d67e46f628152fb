"""Host-side engine of the B200 CoAlign hot path: weight packing (BN folding, K-major bf16), HBM buffer
plan (PF / PS layouts, see include/coalign_b200.h), K-step tables for the implicit-GEMM convolutions and
the launch sequence - captured once per batch signature into a CUDA graph.

PyTorch is used for device memory, streams and graph capture only; every kernel is ours
(libcoalign_b200.so, called through the C ABI).  Mirrors the data flow of
/root/reference/opencood/models/point_pillar_baseline_multiscale.py:93-135 (inference / eval-mode BN).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import CB_OUT_HEADS, CB_OUT_PF, CB_OUT_PS, CB_OUT_UPSAMPLE, ConvDesc

BF16 = torch.bfloat16


def _half_up(x: int) -> int:          # output size of a k3/s2/p1 (and k1/s2/p0) convolution
    return (x + 1) // 2


class Act:
    """One activation buffer in HBM.  layout 'pf' = padded flat, 'ps' = phase split (4 parity planes)."""

    def __init__(self, n_cap: int, H: int, W: int, C_: int, layout: str, precise: bool, device, dtype=BF16, pad: int = 1):
        self.n_cap, self.H, self.W, self.C, self.layout, self.pad = n_cap, H, W, C_, layout, pad
        if layout == "pf":
            self.Hp, self.Wp = H + 2 * pad, W + 2 * pad
            self.plane_rows = 0
            self.rows = n_cap * self.Hp * self.Wp
        else:
            self.Hp, self.Wp = _half_up(H) + 2 * pad, _half_up(W) + 2 * pad      # padded dims of each parity plane
            self.plane_rows = n_cap * self.Hp * self.Wp
            self.rows = 4 * self.plane_rows
        self.precise = precise
        # dtype other than bf16: CPU plan checks only (tests/train_plan_interpreter.py runs the plan in exact fp32)
        self.t = torch.zeros(self.rows * (2 if precise else 1), C_, dtype=dtype, device=device)
        self.lo_off = self.rows * C_ if precise else 0                 # element offset of the lo plane
        self.lo_rows = self.rows if precise else 0

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()

    @property
    def tma_rows(self) -> int:
        return self.rows * (2 if self.precise else 1)

    def zero_(self):
        self.t.zero_()


def bn_fold(sd, prefix: str, eps: float):
    g, b = sd[prefix + ".weight"].double(), sd[prefix + ".bias"].double()
    m, v = sd[prefix + ".running_mean"].double(), sd[prefix + ".running_var"].double()
    s = g / torch.sqrt(v + eps)
    return s, b - m * s


class PackedConv:
    """Weights of one GEMM: bf16 [rows][K] K-major (+ lo part behind it in precise mode), fp32 bias."""

    def __init__(self, w_rows_k: torch.Tensor, bias: torch.Tensor, precise: bool, device, pad_rows_to: int = 1):
        w = w_rows_k.double()
        rows, k = w.shape
        rpad = (rows + pad_rows_to - 1) // pad_rows_to * pad_rows_to
        if rpad != rows:
            w = torch.cat([w, torch.zeros(rpad - rows, k, dtype=w.dtype)], 0)
            bias = torch.cat([bias.double(), torch.zeros(rpad - rows, dtype=torch.float64)], 0)
        w32 = w.float()
        hi = w32.to(BF16)
        if precise:
            lo = (w32 - hi.float()).to(BF16)
            packed = torch.cat([hi, lo], dim=1)
        else:
            packed = hi
        self.w = packed.contiguous().to(device)
        self.bias = bias.float().contiguous().to(device)
        self.rows, self.k = rpad, k
        self.k_total = packed.shape[1]


def pack_conv_weight(w: torch.Tensor, scale: Optional[torch.Tensor]) -> torch.Tensor:
    """[Cout,Cin,kh,kw] (* per-Cout BN scale) -> [Cout, kh*kw*Cin] with K ordered (tap, channel)."""
    w = w.double()
    if scale is not None:
        w = w * scale.view(-1, 1, 1, 1)
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


class CoAlignEngine:
    def __init__(self, args: dict, state_dict: Dict[str, torch.Tensor], max_agents: int, max_scenes: int,
                 device="cuda", precise: bool = False, max_cav: int = 5, block_n_cap: int = 128,
                 use_graph: bool = True, simt_conv: bool = False, pair: bool = True, plan_only: bool = False,
                 backbone: str = "resnet", fusion: bool = True, chan_major: bool = True, pair_min_bn: int = 256,
                 chan_major_256: int = 0,
                 halo: bool = True):
        # backbone "resnet": ResNetBEVBackbone (CoAlign); "plain": BaseBEVBackbone conv stacks (single-agent point_pillar,
        # /root/reference/opencood/models/sub_modules/base_bev_backbone.py).  fusion=False: every agent is its own
        # scene and the per-level maps go straight to the deblocks (point_pillar.py:52-84).
        # plan_only: build weights, buffers and launch descriptors on any torch device WITHOUT loading the CUDA library;
        # nothing can be launched.  Used by the CPU test that interprets the launch plan (tests/plan_interpreter.py).
        self.plan_only = bool(plan_only)
        self.lib = None if plan_only else _lib.load(check_device=True)
        self.args = args
        self.device = torch.device(device)
        self.precise = bool(precise)
        self.max_agents, self.max_scenes, self.max_cav = int(max_agents), int(max_scenes), int(max_cav)
        self.block_n_cap = int(block_n_cap)
        self.use_graph = use_graph
        self.simt_conv = simt_conv            # validation only: evaluate the descriptors with the SIMT kernel
        self.pair = pair                      # CTA-pair (cta_group::2) conv kernel ...
        self.chan_major = bool(chan_major)                                # Cout=128 layers: channel-major 128x256 tiles
        self.chan_major_256 = int(chan_major_256)                         # Cout=256: 0 CTA-pair kernel (default: equal speed, measured), 1 channel-major for 3x3, 2 for all
        self.pair_min_bn = int(pair_min_bn)                               # ... for tiles at least this wide (measured)
        self.halo = bool(halo)                                            # halo-box kernels for the Cout=64/128 layers
        nx, ny, nz = [int(v) for v in args["point_pillar_scatter"]["grid_size"]]
        if nz != 1:
            raise ValueError("PointPillarScatter requires nz == 1")
        self.nx, self.ny = nx, ny
        bb = args["base_bev_backbone"]
        if backbone not in ("resnet", "plain"):
            raise ValueError("backbone must be 'resnet' or 'plain'")
        self.backbone_kind = backbone
        self.fusion = bool(fusion)
        self.layer_nums = list(bb["layer_nums"])
        self.layer_strides = list(bb["layer_strides"])
        self.num_filters = list(bb["num_filters"])
        self.up_strides = list(bb["upsample_strides"])
        self.up_filters = list(bb["num_upsample_filter"])
        self.inplanes = bb.get("inplanes", 64)
        self.c_pfn = args["pillar_vfe"]["num_filters"][-1]
        if self.c_pfn != 64 or self.inplanes != 64 or args["point_pillar_scatter"]["num_features"] != 64:
            raise NotImplementedError("pillar feature width must be 64")
        if any(s not in (1, 2) for s in self.layer_strides):
            raise NotImplementedError("layer strides must be 1 or 2")
        if "compression" in args and args["compression"]:
            raise NotImplementedError("naive compressor is not on the CoAlign path")
        self.method = {"att": 0, "max": 1}[args.get("fusion_method", "att")] if self.fusion else 0
        self.voxel_size = [float(v) for v in args["voxel_size"]]
        self.lidar_range = [float(v) for v in args["lidar_range"]]
        # level geometry
        self.levels: List[Tuple[int, int, int]] = []
        h, w = ny, nx
        for s, c in zip(self.layer_strides, self.num_filters):
            if s == 2:
                h, w = _half_up(h), _half_up(w)
            self.levels.append((h, w, c))
        H0, W0, _ = self.levels[0]
        for (h, w, _c), k in zip(self.levels, self.up_strides):
            if h * k != H0 or w * k != W0:
                raise ValueError("deblock outputs do not line up (torch.cat would fail in the reference too)")
        self._pack_weights(state_dict)
        self._alloc()
        self._graphs: "OrderedDict[tuple, dict]" = OrderedDict()       # LRU, bounded by max_graphs
        self.max_graphs = 16
        self._stream = None if self.plan_only else torch.cuda.Stream(device=self.device)

    # ------------------------------------------------------------------ weights
    def _pack_weights(self, sd):
        sd = {k: v.detach().cpu() for k, v in sd.items()}
        dev, pr = self.device, self.precise
        s, t = bn_fold(sd, "pillar_vfe.pfn_layers.0.norm", 1e-3)
        self.pfn_w = sd["pillar_vfe.pfn_layers.0.linear.weight"].float().contiguous().to(dev)
        self.pfn_scale = s.float().contiguous().to(dev)
        self.pfn_shift = t.float().contiguous().to(dev)
        vs, rg = self.voxel_size, self.lidar_range
        self._vsize_f = np.asarray(vs, np.float32)
        self._range_f = np.asarray(rg, np.float32)
        self._center_off_f = np.asarray([vs[0] / 2 + rg[0], vs[1] / 2 + rg[1], vs[2] / 2 + rg[2]], np.float32)
        self._grid_i = np.asarray([self.nx, self.ny, 1], np.int32)
        self.blocks = []
        inpl = self.inplanes
        for li, (nb, st, pl) in enumerate(zip(self.layer_nums, self.layer_strides, self.num_filters)):
            if self.backbone_kind == "plain":
                # BaseBEVBackbone.blocks[li]: [ZeroPad2d(1), Conv3x3(s, pad 0), BN(eps 1e-3), ReLU] + nb x [Conv3x3, BN, ReLU]
                for j in range(nb + 1):
                    sc, sh = bn_fold(sd, f"backbone.blocks.{li}.{2 + 3 * j}", 1e-3)
                    w = pack_conv_weight(sd[f"backbone.blocks.{li}.{1 + 3 * j}.weight"], sc)
                    self.blocks.append({"layer": li, "k": j, "stride": st if j == 0 else 1, "cin": inpl if j == 0 else pl,
                                        "cout": pl, "pc": PackedConv(w, sh, pr, dev)})
                inpl = pl
                continue
            for k in range(nb):
                p = f"backbone.resnet.layer{li}.{k}"
                s1, t1 = bn_fold(sd, p + ".bn1", 1e-5)
                s2, t2 = bn_fold(sd, p + ".bn2", 1e-5)
                w1 = pack_conv_weight(sd[p + ".conv1.weight"], s1)
                w2 = pack_conv_weight(sd[p + ".conv2.weight"], s2)
                has_ds = (p + ".downsample.0.weight") in sd
                b2 = t2
                if has_ds:
                    sdn, tdn = bn_fold(sd, p + ".downsample.1", 1e-5)
                    wd = pack_conv_weight(sd[p + ".downsample.0.weight"], sdn)
                    w2 = torch.cat([w2, wd], dim=1)                 # extra K blocks: identity branch
                    b2 = t2 + tdn
                self.blocks.append({
                    "layer": li, "k": k, "stride": st if k == 0 else 1, "cin": inpl if k == 0 else pl, "cout": pl,
                    "has_ds": has_ds, "c1": PackedConv(w1, t1, pr, dev), "c2": PackedConv(w2, b2, pr, dev)})
            inpl = pl
        self.deconvs = []
        for i, (k, cu) in enumerate(zip(self.up_strides, self.up_filters)):
            s, t = bn_fold(sd, f"backbone.deblocks.{i}.1", 1e-3)
            w = sd[f"backbone.deblocks.{i}.0.weight"].double() * s.view(1, -1, 1, 1)     # [Cin,Cout,k,k]
            # rows ordered (a, b, co); K = ci
            wp = w.permute(2, 3, 1, 0).reshape(k * k * cu, w.shape[0])
            self.deconvs.append({"k": k, "cout": cu, "cin": w.shape[0], "pc": PackedConv(wp, t, pr, dev)})
        self.c_cat = sum(self.up_filters)
        self.shrink = []
        c_last = self.c_cat
        if "shrink_header" in self.args:
            sh = self.args["shrink_header"]
            for li, (ks, st, pd, dim) in enumerate(zip(sh["kernal_size"], sh["stride"], sh["padding"], sh["dim"])):
                if ks != 3 or st != 1 or pd != 1:
                    raise NotImplementedError("shrink header: only 3x3/s1/p1 is on the B200 path")
                p = f"shrink_conv.layers.{li}.double_conv"
                for idx in (".0", ".2"):
                    w = sd[p + idx + ".weight"]
                    self.shrink.append({"cin": w.shape[1], "cout": w.shape[0],
                                        "pc": PackedConv(pack_conv_weight(w, None), sd[p + idx + ".bias"].double(), pr, dev)})
                c_last = dim
        self.c_last = c_last
        names = (["cls_head", "reg_head"] + (["unc_head"] if "unc_head.weight" in sd else []) +
                 (["dir_head"] if "dir_head.weight" in sd else []))
        hw = torch.cat([sd[n + ".weight"].double().reshape(sd[n + ".weight"].shape[0], -1) for n in names], 0)
        hb = torch.cat([sd[n + ".bias"].double() for n in names], 0)
        self.head_names = [n.replace("_head", "_preds") for n in names]
        self.head_cn = [sd[n + ".weight"].shape[0] for n in names]
        self.head_pad = (hw.shape[0] + 31) // 32 * 32
        if self.head_pad > 256:
            raise NotImplementedError("too many head channels")
        self.head_pc = PackedConv(hw, hb, pr, dev, pad_rows_to=self.head_pad)

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        dev, pr, NA, NS = self.device, self.precise, self.max_agents, self.max_scenes
        in_layout0 = "ps" if self.layer_strides[0] == 2 else "pf"
        self.canvas = Act(NA, self.ny, self.nx, 64, in_layout0, pr, dev)
        self.lvl = []
        for li, (h, w, c) in enumerate(self.levels):
            last = li == len(self.levels) - 1
            out_layout = "pf" if last or self.layer_strides[li + 1] == 1 else "ps"
            self.lvl.append({
                "tmp": Act(NA, h, w, c, "pf", pr, dev),
                "ping": Act(NA, h, w, c, "pf", pr, dev),
                "pong": Act(NA, h, w, c, "pf", pr, dev),
                "out": Act(NA, h, w, c, out_layout, pr, dev),
                "fused": Act(NS, h, w, c, "pf", pr, dev)})
        H0, W0, _ = self.levels[0]
        self.cat = Act(NS, H0, W0, self.c_cat, "pf", pr, dev)
        self.shrink_bufs = [Act(NS, H0, W0, s["cout"], "pf", pr, dev) for s in self.shrink]
        self.head_out = [torch.zeros(NS, cn, H0, W0, dtype=torch.float32, device=dev) for cn in self.head_cn]
        self.affine = torch.zeros(NS, self.max_cav, 2, 3, dtype=torch.float64, device=dev)
        self.pairwise = torch.zeros(NS, self.max_cav, self.max_cav, 4, 4, dtype=torch.float64, device=dev)
        self.agent_off = torch.zeros(NS + 1, dtype=torch.int32, device=dev)
        self._vox_ws = None
        self._dirty_cap = self.max_agents * self.nx * self.ny      # one slot per canvas cell: cannot overflow
        self.dirty_rows = torch.zeros(self._dirty_cap, dtype=torch.int64, device=dev)
        self.dirty_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.pts_buf = None
        self.pt_off_dev = torch.zeros(self.max_agents + 1, dtype=torch.int32, device=dev)   # per-agent point offsets

    # ------------------------------------------------------------------ descriptors
    def _expand(self, steps, lo_rows, k_hi):
        """bf16 mode: steps as is.  precise: hi*hi + lo*hi + hi*lo."""
        if not self.precise:
            return steps
        out = []
        for (ro, col, wk, sel) in steps:
            out.append((ro, col, wk, sel))
            out.append((ro + lo_rows[sel], col, wk, sel))
            out.append((ro, col, wk + k_hi, sel))
        return out

    def _desc(self, a: Sequence[Optional[Act]], pc: PackedConv, steps, n_img: int, Hp: int, Wp: int, n_total: int,
              block_n: int, cout_mod: int, relu: bool, out: Optional[Act], out_mode: int, residual: Optional[Act] = None,
              out_ch_off: int = 0, up_k: int = 0, heads=None) -> ConvDesc:
        d = ConvDesc()
        lo_rows = [0, 0]
        reg = self.__dict__.setdefault("_by_ptr", {})          # pointer -> owning object (plan interpreter / debugging)
        for obj in list(a) + [out, residual]:
            if obj is not None:
                reg[obj.ptr] = obj
        reg[pc.w.data_ptr()] = pc
        reg[pc.bias.data_ptr()] = pc
        for t, _cn in (heads or []):
            reg[t.data_ptr()] = t
        for i, act in enumerate(a):
            if act is not None:
                d.a_ptr[i] = act.ptr
                d.a_rows[i] = act.tma_rows
                d.a_pitch[i] = act.C
                lo_rows[i] = act.lo_rows
        d.w_ptr = pc.w.data_ptr()
        d.w_rows = pc.rows
        d.w_k_total = pc.k_total
        d.n_img, d.Hp, d.Wp = n_img, Hp, Wp
        d.n_total, d.block_n = n_total, block_n
        d.bias = pc.bias.data_ptr()
        d.cout_mod = cout_mod
        d.relu = 1 if relu else 0
        if residual is not None:
            d.residual = residual.ptr
            d.res_pitch = residual.C
            d.res_lo_off = residual.lo_off
        d.out_mode = out_mode
        if out is not None:
            d.out = out.ptr
            d.out_pitch = out.C
            d.out_ch_off = out_ch_off
            d.out_lo_off = out.lo_off
            d.out_Hp, d.out_Wp = out.Hp, out.Wp
            d.out_plane_rows = out.plane_rows
        d.up_k = up_k
        if heads is not None:
            c0 = 0
            for i, (t, cn) in enumerate(heads):
                d.head_out[i] = t.data_ptr()
                d.head_c0[i] = c0
                d.head_cn[i] = cn
                c0 += cn
            d.n_heads = len(heads)
        steps = self._expand(steps, lo_rows, pc.k)
        if len(steps) > _lib.CB_MAX_KSTEPS:
            raise ValueError(f"{len(steps)} K-steps exceed CB_MAX_KSTEPS")
        d.n_ksteps = len(steps)
        for i, (ro, col, wk, sel) in enumerate(steps):
            d.ksteps[i].row_off = int(ro)
            d.ksteps[i].col = int(col)
            d.ksteps[i].w_k = int(wk)
            d.ksteps[i].a_sel = int(sel)
        return d

    @staticmethod
    def _steps_3x3_s1(cin: int, Wp: int, sel: int = 0, k0: int = 0):
        # order (filter row, channel block, filter column): the three taps of a filter row are consecutive K-steps with
        # row shifts d-1, d, d+1, which the halo kernels serve from one TMA box
        return [((r - 1) * Wp + (s - 1), cb * 64, k0 + (r * 3 + s) * cin + cb * 64, sel)
                for r in range(3) for cb in range(cin // 64) for s in range(3)]

    @staticmethod
    def _steps_3x3_s2(cin: int, src: Act, sel: int = 0, k0: int = 0):
        """Input in PS layout: tap (r,s) of a k3/s2/p1 conv reads parity plane ((r+1)&1,(s+1)&1) shifted by
        -1 row/col for r==0 / s==0 (padded plane coords equal padded output coords)."""
        steps = []
        for r in range(3):
            pr_, dr = ((1, -1), (0, 0), (1, 0))[r]
            for s in range(3):
                pc_, dc = ((1, -1), (0, 0), (1, 0))[s]
                ro = (pr_ * 2 + pc_) * src.plane_rows + dr * src.Wp + dc
                for cb in range(cin // 64):
                    steps.append((ro, cb * 64, k0 + (r * 3 + s) * cin + cb * 64, sel))
        return steps

    @staticmethod
    def _steps_1x1(cin: int, sel: int = 0, k0: int = 0):
        return [(0, cb * 64, k0 + cb * 64, sel) for cb in range(cin // 64)]

    def _bn_for(self, cout: int) -> int:
        return min(cout, self.block_n_cap) if cout >= 64 else cout

    def build_descs(self, n_img: int, n_scenes: int) -> List[Tuple[str, object]]:
        """Launch list for one forward after the canvas has been written: ('conv', desc) / ('fuse', level)."""
        ops: List[Tuple[str, object]] = []
        x = self.canvas
        bi = 0
        for li, nb in enumerate(self.layer_nums):
            L = self.lvl[li]
            h, w, c = self.levels[li]
            if self.backbone_kind == "plain":
                for j in range(nb + 1):
                    blk = self.blocks[bi]
                    bi += 1
                    cin, cout, st = blk["cin"], blk["cout"], blk["stride"]
                    dst = L["out"] if j == nb else (L["tmp"] if j % 2 == 0 else L["ping"])
                    steps = self._steps_3x3_s2(cin, x) if st == 2 else self._steps_3x3_s1(cin, x.Wp)
                    mode = CB_OUT_PS if dst.layout == "ps" else CB_OUT_PF
                    ops.append(("conv", self._desc([x, None], blk["pc"], steps, n_img, L["tmp"].Hp, L["tmp"].Wp, cout,
                                                   self._bn_for(cout), cout, True, dst, mode)))
                    x = dst
                ops.append(("fuse" if self.fusion else "copy", li))
                continue
            for k in range(nb):
                blk = self.blocks[bi]
                bi += 1
                last = k == nb - 1
                cin, cout, st = blk["cin"], blk["cout"], blk["stride"]
                bn = self._bn_for(cout)
                tmp = L["tmp"]
                # conv1 (+bn1+relu) -> tmp
                steps1 = self._steps_3x3_s2(cin, x) if st == 2 else self._steps_3x3_s1(cin, x.Wp)
                ops.append(("conv", self._desc([x, None], blk["c1"], steps1, n_img, tmp.Hp, tmp.Wp, cout, bn, cout, True,
                                               tmp, CB_OUT_PF)))
                # conv2 (+bn2) (+ identity / fused 1x1 downsample) + relu -> dst
                dst = L["out"] if last else (L["ping"] if (k % 2 == 0) else L["pong"])
                steps2 = self._steps_3x3_s1(cout, tmp.Wp)
                a1, res = None, None
                if blk["has_ds"]:
                    steps2 = steps2 + self._steps_1x1(cin, sel=1, k0=9 * cout)   # PS plane (0,0) / PF, shift 0
                    a1 = x
                else:
                    res = x
                mode = CB_OUT_PS if dst.layout == "ps" else CB_OUT_PF
                ops.append(("conv", self._desc([tmp, a1], blk["c2"], steps2, n_img, tmp.Hp, tmp.Wp, cout, bn, cout, True,
                                               dst, mode, residual=res)))
                x = dst
            ops.append(("fuse" if self.fusion else "copy", li))
        # decoder: ConvTranspose(k==s)+BN+ReLU as GEMM with pixel-shuffle store into the concat buffer
        ch = 0
        for li, dc in enumerate(self.deconvs):
            f = self.lvl[li]["fused"]
            if not self.fusion and self.lvl[li]["out"].layout == "pf":
                f = self.lvl[li]["out"]                      # no fusion stage: a PF level output feeds the deblock directly
            k, cu = dc["k"], dc["cout"]
            bn = self._bn_for(cu)
            ops.append(("conv", self._desc([f, None], dc["pc"], self._steps_1x1(dc["cin"]), n_scenes, f.Hp, f.Wp,
                                           k * k * cu, bn, cu, True, self.cat, CB_OUT_UPSAMPLE, out_ch_off=ch, up_k=k)))
            ch += cu
        y = self.cat
        for s, buf in zip(self.shrink, self.shrink_bufs):
            bn = self._bn_for(s["cout"])
            ops.append(("conv", self._desc([y, None], s["pc"], self._steps_3x3_s1(s["cin"], y.Wp), n_scenes, y.Hp, y.Wp,
                                           s["cout"], bn, s["cout"], True, buf, CB_OUT_PF)))
            y = buf
        heads = [(t, cn) for t, cn in zip(self.head_out, self.head_cn)]
        ops.append(("conv", self._desc([y, None], self.head_pc, self._steps_1x1(self.c_last), n_scenes, y.Hp, y.Wp,
                                       self.head_pad, 32 if self.head_pad == 32 else self._bn_for(self.head_pad),
                                       self.head_pad, False, None, CB_OUT_HEADS, heads=heads)))
        return ops

    # ------------------------------------------------------------------ launches
    @staticmethod
    def _has_tap_triples(o) -> bool:
        """True for 3x3/s1 K-step tables (first three K-steps = one filter row: same channels, row shifts d, d+1, d+2):
        the halo kernels pay off there; stride-2 / 1x1 tables (no groups) are faster on the plain kernels (measured)."""
        if o.n_ksteps < 3:
            return False
        k = o.ksteps
        return all(k[i].a_sel == k[0].a_sel and k[i].col == k[0].col and k[i].row_off == k[0].row_off + i
                   for i in (1, 2))

    def _launch_ops(self, ops, n_scenes: int, stream_ptr: int):
        if self.plan_only:
            raise RuntimeError("plan_only engine cannot launch (no CUDA library loaded)")
        lib = self.lib
        for kind, o in ops:
            if kind == "conv":
                if self.simt_conv:
                    _lib.check(lib.cb_conv_gemm_simt(C.byref(o), stream_ptr), "cb_conv_gemm_simt")
                elif (self.halo and not self.precise and self._has_tap_triples(o) and o.n_total == 64
                      and o.cout_mod == 64 and o.n_ksteps <= 10 and o.out_mode in (CB_OUT_PF, CB_OUT_PS)):
                    _lib.check(lib.cb_conv_gemm_halo(C.byref(o), 0, stream_ptr), "cb_conv_gemm_halo")
                elif (self.chan_major and not self.precise and o.cout_mod == o.n_total and o.out_mode in (CB_OUT_PF, CB_OUT_PS)
                      and (o.n_total == 128 or (o.n_total == 256 and (self.chan_major_256 == 2 or (
                          self.chan_major_256 == 1 and self.halo and self._has_tap_triples(o)))))):
                    if self.halo and self._has_tap_triples(o):
                        _lib.check(lib.cb_conv_gemm_t_halo(C.byref(o), 0, stream_ptr), "cb_conv_gemm_t_halo")
                    else:
                        _lib.check(lib.cb_conv_gemm_t(C.byref(o), 0, stream_ptr), "cb_conv_gemm_t")
                elif self.pair and o.block_n >= self.pair_min_bn:
                    _lib.check(lib.cb_conv_gemm_pair(C.byref(o), 0, stream_ptr), "cb_conv_gemm_pair")
                else:
                    _lib.check(lib.cb_conv_gemm(C.byref(o), 0, stream_ptr), "cb_conv_gemm")
            elif kind == "copy":
                src, dst = self.lvl[o]["out"], self.lvl[o]["fused"]
                if src.layout == "ps":                       # PF outputs are consumed in place (build_descs)
                    h, w, c = self.levels[o]
                    _lib.check(lib.cb_ps_to_pf(src.ptr, src.lo_off, src.n_cap, n_scenes, h, w, c, dst.ptr, dst.lo_off,
                                               stream_ptr), "cb_ps_to_pf")
            else:
                li = o
                src, dst = self.lvl[li]["out"], self.lvl[li]["fused"]
                h, w, c = self.levels[li]
                _lib.check(lib.cb_warp_att_fuse(src.ptr, 1 if src.layout == "ps" else 0, src.lo_off, src.n_cap,
                                                self.affine.data_ptr(), self.agent_off.data_ptr(), n_scenes,
                                                self.max_cav, h, w, c, self.method, dst.ptr, dst.lo_off, stream_ptr),
                           "cb_warp_att_fuse")

    def _run_backbone(self, ent, n_scenes: int, stream_ptr: int):
        if not self.fusion:
            self._launch_ops(ent["ops"], n_scenes, stream_ptr)
            return
        _lib.check(self.lib.cb_normalize_affine(self.pairwise.data_ptr(), n_scenes, self.max_cav, self.ny, self.nx,
                                                float(self.voxel_size[0]), self.affine.data_ptr(), stream_ptr),
                   "cb_normalize_affine")
        self._launch_ops(ent["ops"], n_scenes, stream_ptr)

    def _graphed(self, key, record_len: Tuple[int, ...], front=None):
        """Run (front-end +) backbone for this batch signature; captured into a CUDA graph on first use."""
        n_img, n_scenes = sum(record_len), len(record_len)
        ent = self._graphs.get(key)
        if ent is None:
            ent = {"ops": self.build_descs(n_img, n_scenes), "graph": None}
            self._graphs[key] = ent
            while len(self._graphs) > self.max_graphs:        # least recently used signature goes (graph + descriptors)
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
        cur = torch.cuda.current_stream(self.device)

        def run(stream_ptr):
            if front is not None:
                front(stream_ptr)
            self._run_backbone(ent, n_scenes, stream_ptr)

        if not self.use_graph:
            run(cur.cuda_stream)
            return
        if ent["graph"] is None:
            run(cur.cuda_stream)          # warm-up outside capture (function attributes, driver entry point)
            cur.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self._stream):
                run(torch.cuda.current_stream(self.device).cuda_stream)
            ent["graph"] = g
        ent["graph"].replay()

    def _set_scene_meta(self, record_len: Tuple[int, ...], pairwise: torch.Tensor):
        n_scenes = len(record_len)
        if n_scenes > self.max_scenes or sum(record_len) > self.max_agents:
            raise ValueError("batch exceeds the engine capacity (max_scenes / max_agents)")
        if not self.fusion:
            if any(v != 1 for v in record_len):
                raise ValueError("a no-fusion engine takes one agent per scene")
            return
        if max(record_len) > min(self.max_cav, 8) or min(record_len) < 1:
            raise ValueError("record_len entries must be in [1, max_cav]")
        if tuple(pairwise.shape[1:]) != (self.max_cav, self.max_cav, 4, 4) or pairwise.shape[0] != n_scenes:
            raise ValueError("pairwise_t_matrix must be (B, max_cav, max_cav, 4, 4)")
        self.pairwise[:n_scenes].copy_(pairwise.to(torch.float64), non_blocking=True)
        off = np.zeros(self.max_scenes + 1, np.int32)
        off[1:n_scenes + 1] = np.cumsum(record_len)
        off[n_scenes + 1:] = off[n_scenes]
        key = tuple(record_len)
        if getattr(self, "_off_key", None) != key:
            if self.plan_only or self.max_scenes + 1 > _lib.CB_MAX_AGENTS + 1:
                self.agent_off.copy_(torch.from_numpy(off), non_blocking=False)
            else:
                _lib.check(self.lib.cb_upload_i32(off.ctypes.data, self.max_scenes + 1, self.agent_off.data_ptr(),
                                                  torch.cuda.current_stream(self.device).cuda_stream), "cb_upload_i32")
            self._off_key = key

    def _outputs(self, n_scenes: int, clone: bool):
        out = {}
        for name, t in zip(self.head_names, self.head_out):
            v = t[:n_scenes]
            out[name] = v.clone() if clone else v
        return out

    # ------------------------------------------------------------------ public forward paths
    @torch.no_grad()
    def forward_voxels(self, voxel_features, voxel_coords, voxel_num_points, record_len: Sequence[int], pairwise,
                       clone: bool = True):
        """Reference input schema: voxel tensors as produced by the dataloader collate."""
        record_len = tuple(int(v) for v in record_len)
        self._set_scene_meta(record_len, pairwise)
        n_img = sum(record_len)
        vf = voxel_features.contiguous()
        vc = voxel_coords.contiguous()
        vn = voxel_num_points.contiguous()
        if vf.dtype != torch.float32 or vc.dtype != torch.int32 or vn.dtype != torch.int32:
            raise TypeError("voxel_features f32, voxel_coords i32, voxel_num_points i32 expected")
        if vf.dim() != 3 or vf.shape[2] != 4 or vf.shape[1] > 32:
            raise ValueError("voxel_features must be (M, max_pts<=32, 4)")
        sp = torch.cuda.current_stream(self.device).cuda_stream
        track = vf.shape[0] <= self._dirty_cap
        self._clear_canvas(sp)
        _lib.check(self.lib.cb_pfn_scatter(vf.data_ptr(), vc.data_ptr(), vn.data_ptr(), vf.shape[0], None,
                                           vf.shape[1], self.pfn_w.data_ptr(), self.pfn_scale.data_ptr(),
                                           self.pfn_shift.data_ptr(), self._vsize_f.ctypes.data,
                                           self._center_off_f.ctypes.data, n_img, self.canvas.n_cap, self.ny, self.nx,
                                           self.canvas.ptr, self.canvas.lo_off,
                                           self.dirty_rows.data_ptr() if track else None,
                                           self.dirty_count.data_ptr() if track else None, sp), "cb_pfn_scatter")
        if not track:
            self._canvas_untracked = True
        self._graphed(("bb", record_len), record_len)
        return self._outputs(len(record_len), clone)

    def _clear_canvas(self, stream_ptr: int):
        """Zero what the previous frame wrote: sparse (dirty-row list) when tracked, full memset otherwise."""
        if getattr(self, "_canvas_untracked", False):
            self.canvas.zero_()
            self.dirty_count.zero_()
            self._canvas_untracked = False
            return
        _lib.check(self.lib.cb_canvas_clear(self.canvas.ptr, self.canvas.lo_off, self.dirty_rows.data_ptr(),
                                            self.dirty_count.data_ptr(), self._dirty_cap, stream_ptr), "cb_canvas_clear")

    def _ws(self, n_agents: int, sum_points: int, max_voxels: int):
        need = self.lib.cb_voxelize_workspace_bytes(n_agents, sum_points, self._grid_i.ctypes.data, max_voxels)
        if self._vox_ws is None or self._vox_ws.numel() < need:
            self._vox_ws = torch.empty(int(need), dtype=torch.uint8, device=self.device)
        return self._vox_ws

    @torch.no_grad()
    def forward_points(self, points, pt_offset: Sequence[int], record_len: Sequence[int], pairwise,
                       max_pts: int = 32, max_voxels: int = 70000, clone: bool = True):
        """Raw clouds (sum_P,4) f32 on the device, agent a = rows pt_offset[a]:pt_offset[a+1]: voxelisation,
        PFN and scatter are fused (no (M,32,4) tensor)."""
        record_len = tuple(int(v) for v in record_len)
        self._set_scene_meta(record_len, pairwise)
        n_img = sum(record_len)
        po = np.ascontiguousarray(pt_offset, dtype=np.int32)
        if po.shape[0] != n_img + 1:
            raise ValueError("pt_offset must have sum(record_len)+1 entries")
        total = int(po[-1])
        if points.dtype != torch.float32 or points.dim() != 2 or points.shape[1] != 4 or points.shape[0] < total:
            raise TypeError("points must be float32 (sum_P, 4)")
        if self.pts_buf is None or self.pts_buf.shape[0] < total:
            # capacity = the caller's buffer (a serving loop passes its largest-frame staging buffer): no regrowth later
            self.pts_buf = torch.empty(max(total, int(points.shape[0]), 1), 4, dtype=torch.float32, device=self.device)
            for k in [k for k in self._graphs if k[0] == "pts"]:
                del self._graphs[k]
        if points.data_ptr() != self.pts_buf.data_ptr():
            self.pts_buf[:total].copy_(points[:total], non_blocking=True)     # static address for the graph
        if po[0] != 0 or (np.diff(po) < 0).any():
            raise ValueError("pt_offset must start at 0 and be non-decreasing")
        # The per-agent offsets live in a device array the kernels read, and the launch geometry depends only on the two
        # capacities below: one captured graph serves every frame of this batch signature, whatever the clouds' sizes
        # (real sweeps differ from frame to frame; a key on the offsets themselves would re-capture every step).
        pt_cap = int(self.pts_buf.shape[0])
        agent_cap = min(pt_cap, (int(np.diff(po).max(initial=0)) + 16383) // 16384 * 16384)
        ws = self._ws(n_img, pt_cap, max_voxels)
        cur = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.cb_upload_i32(po.ctypes.data, n_img + 1, self.pt_off_dev.data_ptr(), cur), "cb_upload_i32")
        if getattr(self, "_canvas_untracked", False):
            self._clear_canvas(cur)                                           # full memset, outside the graph

        def front(stream_ptr):
            self._clear_canvas(stream_ptr)
            _lib.check(self.lib.cb_points_to_canvas_dev(self.pts_buf.data_ptr(), self.pt_off_dev.data_ptr(), n_img,
                                                        pt_cap, agent_cap,
                                                        self._range_f.ctypes.data, self._vsize_f.ctypes.data,
                                                        self._grid_i.ctypes.data, max_pts, max_voxels,
                                                        self.pfn_w.data_ptr(), self.pfn_scale.data_ptr(),
                                                        self.pfn_shift.data_ptr(), self._center_off_f.ctypes.data,
                                                        self.canvas.n_cap, self.canvas.ptr, self.canvas.lo_off,
                                                        self.dirty_rows.data_ptr(), self.dirty_count.data_ptr(),
                                                        ws.data_ptr(), ws.numel(), stream_ptr), "cb_points_to_canvas_dev")

        key = ("pts", record_len, pt_cap, agent_cap, int(max_pts), int(max_voxels), ws.data_ptr())
        self._graphed(key, record_len, front)
        return self._outputs(len(record_len), clone)

    @torch.no_grad()
    def run_front_only(self, pt_offset: Sequence[int], max_pts: int = 32, max_voxels: int = 70000):
        """Measurement hook: the pillar front-end alone (canvas clear + voxelise + PFN + scatter) on `pts_buf`."""
        po = np.ascontiguousarray(pt_offset, dtype=np.int32)
        n_img = po.shape[0] - 1
        ws = self._ws(n_img, int(po[-1]), max_voxels)
        sp = torch.cuda.current_stream(self.device).cuda_stream
        self._clear_canvas(sp)
        _lib.check(self.lib.cb_points_to_canvas(self.pts_buf.data_ptr(), po.ctypes.data, n_img,
                                                self._range_f.ctypes.data, self._vsize_f.ctypes.data,
                                                self._grid_i.ctypes.data, max_pts, max_voxels,
                                                self.pfn_w.data_ptr(), self.pfn_scale.data_ptr(),
                                                self.pfn_shift.data_ptr(), self._center_off_f.ctypes.data,
                                                self.canvas.n_cap, self.canvas.ptr, self.canvas.lo_off,
                                                self.dirty_rows.data_ptr(), self.dirty_count.data_ptr(),
                                                ws.data_ptr(), ws.numel(), sp), "cb_points_to_canvas")

    @torch.no_grad()
    def voxelize(self, points, pt_offset: Sequence[int], max_pts: int = 32, max_voxels: int = 70000):
        """A1/A2 on the GPU, reference output format; returns (voxels, coords[a,z,y,x], num_points) trimmed."""
        po = np.asarray(pt_offset, np.int32)
        n_agents = po.shape[0] - 1
        total = int(po[-1])
        cap = max(1, min(total, n_agents * max_voxels))
        ws = self._ws(n_agents, total, max_voxels)
        vox = torch.empty(cap, max_pts, 4, dtype=torch.float32, device=self.device)
        crd = torch.empty(cap, 4, dtype=torch.int32, device=self.device)
        npt = torch.empty(cap, dtype=torch.int32, device=self.device)
        nv = torch.zeros(n_agents + 1, dtype=torch.int32, device=self.device)
        sp = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.cb_voxelize(points.contiguous().data_ptr(), po.ctypes.data, n_agents,
                                        self._range_f.ctypes.data, self._vsize_f.ctypes.data, self._grid_i.ctypes.data,
                                        max_pts, max_voxels, vox.data_ptr(), crd.data_ptr(), npt.data_ptr(),
                                        nv.data_ptr(), ws.data_ptr(), ws.numel(), sp), "cb_voxelize")
        m = int(nv[-1].item())
        return vox[:m], crd[:m], npt[:m], nv

    # ------------------------------------------------------------------ debugging / tests
    def read_act(self, act: Act, n: int, ch_off: int = 0, channels: Optional[int] = None) -> torch.Tensor:
        """Activation buffer -> dense NCHW float32 (hi+lo)."""
        c = channels if channels is not None else act.C
        n_read = act.n_cap if act.layout == "ps" else n          # PS plane stride is the buffer capacity
        out = torch.empty(n_read, c, act.H, act.W, dtype=torch.float32, device=self.device)
        sp = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.cb_layout_to_nchw(act.ptr, act.lo_off, 1 if act.layout == "ps" else 0, n_read, c, act.H,
                                              act.W, act.C, ch_off, out.data_ptr(), sp), "cb_layout_to_nchw")
        return out[:n]
