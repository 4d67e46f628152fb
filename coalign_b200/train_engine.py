"""Host side of the device training step: parameter / gradient flat buffers, per-step weight packing,
HBM buffer plan and the launch lists of the train-mode forward and of the backward pass.

What runs where (all kernels ours, through the C ABI of include/coalign_b200.h "Training step, device side"):
  forward   conv GEMMs with raw weights (cb_conv_gemm*) -> cb_bn_stats / cb_bn_finalize / cb_bn_apply (batch statistics,
            running-stat update), cb_pfn_train_* + cb_pfn_scatter, cb_warp_att_fuse, heads GEMM
  backward  cb_heads_grad_pack, cb_bn_bwd_reduce / cb_bn_bwd_apply, dgrad = the forward GEMM kernel on dZ with
            transposed / flipped packed weights, wgrad = cb_wgrad (MN-major tcgen05 split-K), cb_warp_att_fuse_bwd +
            cb_grad_combine, cb_pfn_bwd
Mirrors autograd of /root/reference/opencood/models/point_pillar_baseline_multiscale.py:93-135 in `.train()` mode (the loop
body of tools/train.py:105-125); the oracle of every piece is oracle/backward_oracle.py.

The launch lists are plain python lists of (kind, dict) so that tests/train_plan_interpreter.py can execute the same plan on
the CPU (descriptor tables, layouts, maps and unit lists are host logic; checked without a GPU).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import CB_OUT_HEADS, CB_OUT_PF, CB_OUT_PS, CB_OUT_UPSAMPLE, Map, WgradDesc
from .engine import BF16, Act, CoAlignEngine, _half_up


class ActView:
    """One parity plane of a PS activation seen as a PF tensor (the output of a stride-2 dgrad GEMM)."""

    def __init__(self, parent: Act, plane: int):
        self.parent, self.plane = parent, plane
        self.C, self.Hp, self.Wp, self.H, self.W = parent.C, parent.Hp, parent.Wp, parent.Hp - 2, parent.Wp - 2
        self.layout, self.plane_rows, self.n_cap = "pf", 0, parent.n_cap
        self.row0 = plane * parent.plane_rows
        self.rows = parent.plane_rows
        self.precise, self.lo_off, self.lo_rows = parent.precise, parent.lo_off, parent.lo_rows
        self.t = parent.t

    @property
    def ptr(self) -> int:
        return self.parent.ptr + self.row0 * self.C * 2

    @property
    def tma_rows(self) -> int:
        return self.parent.tma_rows - self.row0


class TrainPack:
    """bf16 packed weights of one GEMM, refreshed from the fp32 parameters every step (cb_pack_weight)."""

    def __init__(self, rows: int, k: int, precise: bool, device, bias: Optional[torch.Tensor] = None, pad_rows_to: int = 1,
                 dtype=BF16):
        self.rows = (rows + pad_rows_to - 1) // pad_rows_to * pad_rows_to
        self.k = k
        self.k_total = k * (2 if precise else 1)
        self.w = torch.zeros(self.rows, self.k_total, dtype=dtype, device=device)
        self.bias = bias if bias is not None else torch.zeros(self.rows, dtype=torch.float32, device=device)
        self.jobs: List[tuple] = []          # (param view, R1, R0, K1, K0, s_r1, s_r0, s_k1, s_k0, row offset, k offset)

    def add(self, src: torch.Tensor, R1, R0, K1, K0, s_r1, s_r0, s_k1, s_k0, row_off=0, k_off=0):
        self.jobs.append((src, R1, R0, K1, K0, s_r1, s_r0, s_k1, s_k0, row_off, k_off))


class PackRows:
    """Rows [r0, r0 + n) of a TrainPack as a GEMM weight matrix of its own (a GEMM whose N exceeds what one launch's
    epilogue supports is split into column blocks)."""

    def __init__(self, pk: TrainPack, r0: int, n: int):
        self.w, self.bias = pk.w[r0:r0 + n], pk.bias[r0:r0 + n]
        self.rows, self.k, self.k_total = n, pk.k, pk.k_total


class TrainEngine(CoAlignEngine):
    """Training-mode engine.  `state_dict` provides the initial parameters and BatchNorm running statistics; afterwards the
    engine owns them in two flat fp32 buffers (`pflat`, `gflat` for the gradients, same layout, ordered in the order the
    backward pass completes them so that gradient buckets are contiguous slices)."""

    def __init__(self, args: dict, state_dict: Dict[str, torch.Tensor], max_agents: int, max_scenes: int, device="cuda",
                 precise: bool = False, max_cav: int = 5, block_n_cap: int = 256, plan_only: bool = False,
                 max_voxels_total: int = 0, max_pts: int = 32, exact_fp32_plan: bool = False):
        # exact_fp32_plan: plan_only engines may keep activations / packed weights in float32, so that the CPU interpreter
        # checks the launch plan against the fp32 oracle without bf16 rounding in between (never on the GPU)
        if exact_fp32_plan and not plan_only:
            raise ValueError("exact_fp32_plan is a CPU plan-check option")
        self.act_dtype = torch.float32 if exact_fp32_plan else BF16
        self.max_voxels_total = int(max_voxels_total) if max_voxels_total else int(max_agents) * 32000
        self.max_pts = int(max_pts)
        super().__init__(args, state_dict, max_agents, max_scenes, device=device, precise=precise, max_cav=max_cav,
                         block_n_cap=block_n_cap, use_graph=False, plan_only=plan_only)
        if self.backbone_kind != "resnet" or not self.fusion:
            raise NotImplementedError("training is implemented for the CoAlign model (ResNet backbone + fusion)")
        if any(s != 2 for s in self.layer_strides):
            raise NotImplementedError("training path: every level has stride 2 (all CoAlign yamls)")
        h, w = self.ny, self.nx
        for _ in self.layer_strides:
            if h % 2 or w % 2:
                raise NotImplementedError("training path needs even map sizes at every stride-2 level")
            h, w = h // 2, w // 2

    # ------------------------------------------------------------------ parameters
    def _param_order(self, sd) -> List[str]:
        names = []
        heads = ["cls_head", "reg_head"] + (["dir_head"] if "dir_head.weight" in sd else [])
        names += [h + ".weight" for h in heads] + [h + ".bias" for h in heads]
        if "shrink_header" in self.args:
            for li in reversed(range(len(self.args["shrink_header"]["kernal_size"]))):
                p = f"shrink_conv.layers.{li}.double_conv"
                names += [p + ".2.weight", p + ".2.bias", p + ".0.weight", p + ".0.bias"]
        for i in reversed(range(len(self.up_strides))):
            names += [f"backbone.deblocks.{i}.0.weight", f"backbone.deblocks.{i}.1.weight", f"backbone.deblocks.{i}.1.bias"]
        self._bucket_marks = [len(names)]
        for li in reversed(range(len(self.layer_nums))):
            for k in reversed(range(self.layer_nums[li])):
                p = f"backbone.resnet.layer{li}.{k}"
                names += [p + ".conv2.weight", p + ".bn2.weight", p + ".bn2.bias"]
                if p + ".downsample.0.weight" in sd:
                    names += [p + ".downsample.0.weight", p + ".downsample.1.weight", p + ".downsample.1.bias"]
                names += [p + ".conv1.weight", p + ".bn1.weight", p + ".bn1.bias"]
            self._bucket_marks.append(len(names))
        names += ["pillar_vfe.pfn_layers.0.linear.weight", "pillar_vfe.pfn_layers.0.norm.weight",
                  "pillar_vfe.pfn_layers.0.norm.bias"]
        return names

    def _pack_weights(self, sd):
        sd = {k: v.detach() for k, v in sd.items()}
        dev, pr = self.device, self.precise
        order = self._param_order(sd)
        learn = {k for k, v in sd.items() if v.is_floating_point() and "running_" not in k}
        if set(order) != learn:
            raise ValueError(f"unexpected parameter set: {sorted(set(order) ^ learn)[:6]}")
        self.param_names = order
        offs, n = {}, 0
        for name in order:
            offs[name] = n
            n += (sd[name].numel() + 3) // 4 * 4                  # 16-byte aligned slots (vector reductions, TMA)
        self.param_off, self.n_flat = offs, n
        self.pflat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.gflat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.P: Dict[str, torch.Tensor] = {}
        self.G: Dict[str, torch.Tensor] = {}
        for name in order:
            o, t = offs[name], sd[name]
            self.P[name] = self.pflat[o:o + t.numel()].view(t.shape)
            self.G[name] = self.gflat[o:o + t.numel()].view(t.shape)
            self.P[name].copy_(t.float())
        marks = self._bucket_marks + [len(order)]
        ends = [offs[order[m]] if m < len(order) else n for m in marks]
        self.buckets = [(a, b) for a, b in zip([0] + ends[:-1], ends) if b > a]   # gradient slices in completion order
        # running statistics (buffers)
        self.bn_names = sorted({k[:-len(".running_mean")] for k in sd if k.endswith(".running_mean")})
        self.R: Dict[str, torch.Tensor] = {}
        for b in self.bn_names:
            self.R[b + ".running_mean"] = sd[b + ".running_mean"].float().clone().to(dev)
            self.R[b + ".running_var"] = sd[b + ".running_var"].float().clone().to(dev)
        self.num_batches_tracked = 0
        vs, rg = self.voxel_size, self.lidar_range
        self._vsize_f = np.asarray(vs, np.float32)
        self._range_f = np.asarray(rg, np.float32)
        self._center_off_f = np.asarray([vs[0] / 2 + rg[0], vs[1] / 2 + rg[1], vs[2] / 2 + rg[2]], np.float32)
        self._grid_i = np.asarray([self.nx, self.ny, 1], np.int32)
        # ---- packed weights (forward "f" and input-gradient "d" layouts); K ordered (tap, channel)
        self.packs: List[TrainPack] = []

        def conv_packs(name, cout, cin, taps, extra_d=None):
            w = self.P[name]
            f = TrainPack(cout, taps * cin, pr, dev, dtype=self.act_dtype)
            f.add(w, 1, cout, taps, cin, 0, cin * taps, 1, taps)
            kd = taps * cout + (extra_d[1] if extra_d else 0)
            d = TrainPack(cin, kd, pr, dev, dtype=self.act_dtype)
            d.add(w, 1, cin, taps, cout, 0, taps, 1, cin * taps)
            if extra_d:                                            # 1x1 downsample of the same input: extra K block
                d.add(self.P[extra_d[0]], 1, cin, 1, extra_d[1], 0, 1, 0, cin, 0, taps * cout)
            self.packs += [f, d]
            return f, d

        self.blocks = []
        inpl = self.inplanes
        for li, (nb, st, pl) in enumerate(zip(self.layer_nums, self.layer_strides, self.num_filters)):
            for k in range(nb):
                p = f"backbone.resnet.layer{li}.{k}"
                cin = inpl if k == 0 else pl
                has_ds = (p + ".downsample.0.weight") in sd
                blk = {"layer": li, "k": k, "p": p, "stride": st if k == 0 else 1, "cin": cin, "cout": pl, "has_ds": has_ds}
                blk["f1"], blk["d1"] = conv_packs(p + ".conv1.weight", pl, cin, 9, (p + ".downsample.0.weight", pl) if has_ds else None)
                blk["f2"], blk["d2"] = conv_packs(p + ".conv2.weight", pl, pl, 9)
                if has_ds:
                    fd = TrainPack(pl, cin, pr, dev, dtype=self.act_dtype)
                    fd.add(self.P[p + ".downsample.0.weight"], 1, pl, 1, cin, 0, cin, 0, 1)
                    self.packs.append(fd)
                    blk["fd"] = fd
                self.blocks.append(blk)
            inpl = pl
        self.deconvs = []
        for i, (k, cu) in enumerate(zip(self.up_strides, self.up_filters)):
            name = f"backbone.deblocks.{i}.0.weight"
            cin = self.P[name].shape[0]
            f = TrainPack(k * k * cu, cin, pr, dev, dtype=self.act_dtype)                                   # rows (a, b, co); K = ci
            f.add(self.P[name], k * k, cu, 1, cin, 1, k * k, 0, cu * k * k)
            d = TrainPack(cin, k * k * cu, pr, dev, dtype=self.act_dtype)                                   # rows ci; K = (a, b, co)
            d.add(self.P[name], 1, cin, k * k, cu, 0, cu * k * k, 1, k * k)
            self.packs += [f, d]
            self.deconvs.append({"k": k, "cout": cu, "cin": cin, "f": f, "d": d, "name": name, "bn": f"backbone.deblocks.{i}.1"})
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
                    wname = p + idx + ".weight"
                    cout, cin = self.P[wname].shape[:2]
                    f = TrainPack(cout, 9 * cin, pr, dev, bias=self.P[p + idx + ".bias"], dtype=self.act_dtype)
                    f.add(self.P[wname], 1, cout, 9, cin, 0, cin * 9, 1, 9)
                    d = TrainPack(cin, 9 * cout, pr, dev, dtype=self.act_dtype)
                    d.add(self.P[wname], 1, cin, 9, cout, 0, 9, 1, cin * 9)
                    self.packs += [f, d]
                    self.shrink.append({"cin": cin, "cout": cout, "f": f, "d": d, "name": p + idx})
                c_last = dim
        self.c_last = c_last
        heads = ["cls_head", "reg_head"] + (["dir_head"] if "dir_head.weight" in sd else [])
        self.head_names = [h.replace("_head", "_preds") for h in heads]
        self.head_mods = heads
        self.head_cn = [self.P[h + ".weight"].shape[0] for h in heads]
        tot = sum(self.head_cn)
        self.head_pad = 32
        if tot > 32:
            raise NotImplementedError("training path: at most 32 head channels")
        self.head_bias = torch.zeros(32, dtype=torch.float32, device=dev)              # gathered every step (3 tiny copies)
        self.head_dbias = torch.zeros(32, dtype=torch.float32, device=dev)
        hf = TrainPack(tot, c_last, pr, dev, bias=self.head_bias, pad_rows_to=32, dtype=self.act_dtype)
        hd = TrainPack(c_last, 64, pr, dev, dtype=self.act_dtype)
        c0 = 0
        for h, cn in zip(heads, self.head_cn):
            hf.add(self.P[h + ".weight"], 1, cn, 1, c_last, 0, c_last, 0, 1, c0, 0)
            hd.add(self.P[h + ".weight"], 1, c_last, 1, cn, 0, 1, 0, c_last, 0, c0)
            c0 += cn
        self.packs += [hf, hd]
        self.head_f, self.head_d = hf, hd
        self.pfn_w = self.P["pillar_vfe.pfn_layers.0.linear.weight"]
        # packed-order fp32 gradient scratch of the GEMM weights (unpacked into gflat by cb_permute_f32)
        self.wg_slots, n_wg = {}, 0
        for blk in self.blocks:
            for key, cout, cin, taps in (("conv1", blk["cout"], blk["cin"], 9), ("conv2", blk["cout"], blk["cout"], 9)):
                self.wg_slots[blk["p"] + "." + key + ".weight"] = (n_wg, cout, taps * cin)
                n_wg += cout * taps * cin
            if blk["has_ds"]:
                self.wg_slots[blk["p"] + ".downsample.0.weight"] = (n_wg, blk["cout"], blk["cin"])
                n_wg += blk["cout"] * blk["cin"]
        for dc in self.deconvs:
            self.wg_slots[dc["name"]] = (n_wg, dc["k"] ** 2 * dc["cout"], dc["cin"])
            n_wg += dc["k"] ** 2 * dc["cout"] * dc["cin"]
        for s in self.shrink:
            self.wg_slots[s["name"] + ".weight"] = (n_wg, s["cout"], 9 * s["cin"])
            n_wg += s["cout"] * 9 * s["cin"]
        self.wgflat = torch.zeros(n_wg, dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ buffers
    def _alloc(self):
        dev, pr, NA, NS = self.device, self.precise, self.max_agents, self.max_scenes
        A = lambda n, h, w, c, layout="pf": Act(n, h, w, c, layout, pr, dev, self.act_dtype)      # noqa: E731
        self.canvas = A(NA, self.ny, self.nx, 64, "ps")
        self.d_canvas = A(NA, self.ny, self.nx, 64, "ps")
        self.lvl = []
        bi = 0
        nbn = 0
        for li, (h, w, c) in enumerate(self.levels):
            last = li == len(self.levels) - 1
            out_layout = "pf" if last else "ps"
            blocks = []
            for k in range(self.layer_nums[li]):
                blk = self.blocks[bi]
                bi += 1
                lastb = k == self.layer_nums[li] - 1
                blocks.append({"z1": A(NA, h, w, c), "o1": A(NA, h, w, c), "z2": A(NA, h, w, c),
                               "zd": A(NA, h, w, c) if blk["has_ds"] else None,
                               "out": A(NA, h, w, c, out_layout if lastb else "pf")})
            self.lvl.append({
                "blocks": blocks, "out": blocks[-1]["out"], "fused": A(NS, h, w, c),
                "d_out": A(NA, h, w, c, out_layout), "ga": A(NA, h, w, c), "gb": A(NA, h, w, c),
                "dz": A(NA, h, w, c), "dzd": A(NA, h, w, c), "d_o1": A(NA, h, w, c), "dsum": A(NA, h, w, c),
                "d_fused": A(NS, h, w, c),
                "dfeat": torch.zeros(NA * h * w * c, dtype=torch.float32, device=dev)})
        H0, W0, _ = self.levels[0]
        self.cat = A(NS, H0, W0, self.c_cat)
        self.d_cat = A(NS, H0, W0, self.c_cat)
        for li, dc in enumerate(self.deconvs):
            h, w, _c = self.levels[li]
            dc["zu"] = A(NS, H0, W0, dc["cout"])                    # z pixel-shuffled to full resolution (CB_OUT_UPSAMPLE)
            dc["dzu"] = A(NS, h, w, dc["k"] ** 2 * dc["cout"])       # dz in the GEMM's row space: columns (a, b, co)
        self.shrink_bufs = [A(NS, H0, W0, s["cout"]) for s in self.shrink]
        self.d_shrink = [A(NS, H0, W0, s["cout"]) for s in self.shrink]
        self.dz_shrink = [A(NS, H0, W0, s["cout"]) for s in self.shrink]
        self.head_out = [torch.zeros(NS, cn, H0, W0, dtype=torch.float32, device=dev) for cn in self.head_cn]
        self.head_grad = [torch.zeros(NS, cn, H0, W0, dtype=torch.float32, device=dev) for cn in self.head_cn]
        self.g_pf = A(NS, H0, W0, 64)
        self.affine = torch.zeros(NS, self.max_cav, 2, 3, dtype=torch.float64, device=dev)
        self.pairwise = torch.zeros(NS, self.max_cav, self.max_cav, 4, 4, dtype=torch.float64, device=dev)
        self.agent_off = torch.zeros(NS + 1, dtype=torch.int32, device=dev)
        # per-BatchNorm fp32 vectors (scale, shift, mean, inv_std) and fp64 reduction slots (forward sums, backward sums)
        self.bn_slot: Dict[str, dict] = {}
        nf, nd = 0, 0
        for b in self.bn_names:
            c = self.R[b + ".running_mean"].numel()
            self.bn_slot[b] = {"c": c, "f": nf, "d": nd // 2}
            nf += 4 * c
            nd += 4 * c
        for s in self.shrink:                                     # bias-only "norms" of the shrink convs: backward sums
            self.bn_slot[s["name"]] = {"c": s["cout"], "f": nf, "d": nd // 2}
            nf += 4 * s["cout"]
            nd += 4 * s["cout"]
        self.bnf = torch.zeros(nf, dtype=torch.float32, device=dev)
        # fp64 reduction slots: forward region (per norm [2c] + the PFN's 65 feature sums) / backward region (per norm [2c]
        # + the PFN's 768 sums); each region is zeroed once at the start of its pass
        self.bn_ticket = torch.zeros(1, dtype=torch.int32, device=dev)       # cb_bn_stats_finalize CTA counter
        self.red_f = torch.zeros(nd // 2 + 128, dtype=torch.float64, device=dev)
        self.red_b = torch.zeros(nd // 2 + 768, dtype=torch.float64, device=dev)
        self.pfn_red_off = nd // 2
        self.pfn_bsum_off = nd // 2
        # voxel tensors of the batch (static capacity: the CUDA graph of a step reads them in place)
        mv = self.max_voxels_total
        self.vox_f = torch.zeros(mv, self.max_pts, 4, dtype=torch.float32, device=dev)
        self.vox_c = torch.zeros(mv, 4, dtype=torch.int32, device=dev)
        self.vox_n = torch.zeros(mv, dtype=torch.int32, device=dev)
        self.n_vox_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.dirty_rows = None
        self.dirty_count = None
        self.pts_buf = None
        self._vox_ws = None

    # ------------------------------------------------------------------ small helpers
    def bnv(self, name: str, which: int) -> torch.Tensor:
        """which: 0 scale, 1 shift, 2 mean, 3 inv_std."""
        s = self.bn_slot[name]
        return self.bnf[s["f"] + which * s["c"]: s["f"] + (which + 1) * s["c"]]

    def redv(self, name: str, which: int) -> torch.Tensor:
        """which: 0 forward sums [2c], 1 backward sums [2c]."""
        s = self.bn_slot[name]
        return (self.red_b if which else self.red_f)[s["d"]: s["d"] + 2 * s["c"]]

    def _map(self, n_img: int, z: Act, y, c_mod: Optional[int] = None, y_ch_off: int = 0, up_k: int = 0,
             z_at_y_pitch: int = 0) -> Map:
        m = Map()
        m.n_img, m.Hp, m.Wp, m.c_total, m.c_mod = n_img, z.Hp, z.Wp, z.C, c_mod or z.C
        m.z_at_y, m.z_pitch = (1, z_at_y_pitch) if z_at_y_pitch else (0, 0)
        if up_k:
            m.y_mode = CB_OUT_UPSAMPLE
        else:
            m.y_mode = CB_OUT_PS if y.layout == "ps" else CB_OUT_PF
        m.y_pitch, m.y_ch_off, m.up_k = y.C, y_ch_off, up_k
        m.y_Hp, m.y_Wp, m.y_plane_rows = y.Hp, y.Wp, y.plane_rows
        return m

    def _steps_dgrad_3x3_s1(self, cout: int, Wp: int):
        # adjoint of _steps_3x3_s1: negated row shifts, K ordered (tap, co block); filter-row triples ascend (halo kernels)
        return [(-((r - 1) * Wp + (s - 1)), cb * 64, (r * 3 + s) * cout + cb * 64, 0)
                for r in range(3) for cb in range(cout // 64) for s in (2, 1, 0)]

    @staticmethod
    def _s2_taps():
        # forward tap r (or s) of a k3/s2/p1 conv reads parity plane p at padded-row shift d (engine._steps_3x3_s2)
        return ((1, -1), (0, 0), (1, 0))

    def _steps_dgrad_3x3_s2(self, cout: int, Wp: int, ph: int, pw: int, ds_cout: int = 0):
        """Input-gradient GEMM of one parity plane (ph, pw) of a k3/s2/p1 conv: the taps that read this plane, shifts
        negated; plane (0,0) also receives the 1x1/s2 downsample branch (A operand 1 = its dZ, weights appended along K)."""
        steps = []
        for r, (pr_, dr) in enumerate(self._s2_taps()):
            if pr_ != ph:
                continue
            for s, (pc_, dc) in enumerate(self._s2_taps()):
                if pc_ != pw:
                    continue
                for cb in range(cout // 64):
                    steps.append((-(dr * Wp + dc), cb * 64, (r * 3 + s) * cout + cb * 64, 0))
        if ph == 0 and pw == 0 and ds_cout:
            steps += [(0, cb * 64, 9 * cout + cb * 64, 1) for cb in range(ds_cout // 64)]
        return steps

    def _wgrad(self, dz: Act, xs: Sequence[Optional[Act]], boxes: List[Tuple[int, int, int, int]], n_img: int, m_total: int,
               slot: str, ld: int, dst: Optional[torch.Tensor] = None) -> List[Tuple[str, dict]]:
        """boxes: (x_sel, row_off, col, out column offset).  One op per <= CB_WGRAD_MAX_UNITS units."""
        if dst is None:
            off, rows, ld_ = self.wg_slots[slot]
            assert rows >= m_total and ld_ == ld, (slot, rows, m_total, ld_, ld)
            dst_t, dst_off = self.wgflat, off
        else:
            dst_t, dst_off = dst, 0
        units = []
        nu = (len(boxes) + _lib.CB_WGRAD_MAX_BOXES - 1) // _lib.CB_WGRAD_MAX_BOXES
        cuts = [len(boxes) * i // nu for i in range(nu + 1)]          # equally sized units (9 boxes -> 3+3+3, not 4+4+1):
        for m0 in range(0, m_total, 128):                              # every work item of the launch takes the same time
            for a, b in zip(cuts[:-1], cuts[1:]):
                units.append((m0, min(128, m_total - m0), boxes[a:b]))
        ops = []
        for u0 in range(0, len(units), _lib.CB_WGRAD_MAX_UNITS):
            d = WgradDesc()
            d.dz_ptr = dz.ptr
            d.dz_lo_ptr = dz.ptr + dz.lo_off * 2 if dz.precise else None
            d.dz_pitch, d.k_splits = dz.C, 0
            d.rows_total = n_img * dz.Hp * dz.Wp
            for i, x in enumerate(xs):
                if x is not None:
                    d.x_ptr[i], d.x_rows[i], d.x_pitch[i], d.x_lo_rows[i] = x.ptr, x.tma_rows, x.C, x.lo_rows
            d.dw = dst_t.data_ptr() + dst_off * 4
            grp = units[u0:u0 + _lib.CB_WGRAD_MAX_UNITS]
            d.n_units = len(grp)
            for ui, (m0, mv, bx) in enumerate(grp):
                u = d.units[ui]
                u.m0, u.a_row_off, u.m_valid, u.n_boxes = m0, 0, mv, len(bx)
                for j, (sel, ro, col, oc) in enumerate(bx):
                    u.box[j].row_off, u.box[j].out_ld, u.box[j].out_off = int(ro), ld, m0 * ld + oc
                    u.box[j].col, u.box[j].x_sel = col, sel
            ops.append(("wgrad", {"desc": d, "dz": dz, "xs": list(xs), "dst": dst_t, "dst_off": dst_off, "units": grp, "ld": ld}))
        return ops

    @staticmethod
    def permute_job_shape(q: dict):
        """(R1, R0, K0, (s_r1, s_r0, s_k1, s_k0)) of a gradient permutation: dst[(r1, r0)][k0] = src[r1*s_r1 + r0*s_r0 + k0*s_k0]
        (the cb_permute_f32 / cb_pack_job convention with K1 = 1)."""
        if q["kind"] == "conv":                       # packed [co][tap][ci] -> OIHW: dst[(co, ci)][tap]
            co, ci, t = q["cout"], q["cin"], q["taps"]
            return co, ci, t, (t * ci, 1, 0, ci)
        ci, cu, k = q["cin"], q["cu"], q["k"]         # packed [(ab, co)][ci] -> [ci][co][ab]: dst[(ci, co)][ab]
        return ci, cu, k * k, (1, ci, 0, cu * ci)

    def _unpack(self, slot: str, kind: str, **kw) -> Tuple[str, dict]:
        """packed-order gradient -> the parameter's own layout in gflat (cb_permute_f32)."""
        off, rows, ld = self.wg_slots[slot]
        return ("permute", {"src": self.wgflat, "src_off": off, "dst": self.G[slot], "kind": kind, **kw})

    # ------------------------------------------------------------------ launch lists
    def build_train_ops(self, record_len: Sequence[int]) -> Dict[str, list]:
        n_img, n_sc = sum(record_len), len(record_len)
        fwd: List[Tuple[str, dict]] = []
        bwd: List[Tuple[str, dict]] = []
        conv = lambda d: ("conv", {"desc": d})        # noqa: E731

        def bn_(c):                                   # widest tile that divides the GEMM's N (384 -> 128)
            for b in (256, 128, 64, 32):
                if b <= max(self.block_n_cap, 64) and c % b == 0:
                    return b
            raise ValueError(f"no tile width divides {c}")
        # ============================================================ forward
        fwd.append(("zero_fwd", {}))
        fwd.append(("pack_all", {}))
        fwd.append(("pfn_fwd", {"n_img": n_img}))
        x = self.canvas
        bi = 0
        for li, nb in enumerate(self.layer_nums):
            L = self.lvl[li]
            h, w, c = self.levels[li]
            cnt = float(n_img * h * w)
            for k in range(nb):
                blk, B = self.blocks[bi], L["blocks"][k]
                bi += 1
                p, cin, cout = blk["p"], blk["cin"], blk["cout"]
                z1, o1, z2, out = B["z1"], B["o1"], B["z2"], B["out"]
                steps1 = self._steps_3x3_s2(cin, x) if blk["stride"] == 2 else self._steps_3x3_s1(cin, x.Wp)
                fwd.append(conv(self._desc([x, None], blk["f1"], steps1, n_img, z1.Hp, z1.Wp, cout, bn_(cout), cout, False,
                                           z1, CB_OUT_PF)))
                m1 = self._map(n_img, z1, o1)
                fwd.append(("bn_fwd", {"bn": p + ".bn1", "z": z1, "map": m1, "count": cnt, "eps": 1e-5, "mom": 0.1}))
                fwd.append(("bn_apply", {"bn": p + ".bn1", "z": z1, "map": m1, "y": o1, "relu": 1}))
                fwd.append(conv(self._desc([o1, None], blk["f2"], self._steps_3x3_s1(cout, o1.Wp), n_img, z2.Hp, z2.Wp, cout,
                                           bn_(cout), cout, False, z2, CB_OUT_PF)))
                m2 = self._map(n_img, z2, out)
                fwd.append(("bn_fwd", {"bn": p + ".bn2", "z": z2, "map": m2, "count": cnt, "eps": 1e-5, "mom": 0.1}))
                if blk["has_ds"]:
                    zd = B["zd"]
                    fwd.append(conv(self._desc([x, None], blk["fd"], self._steps_1x1(cin), n_img, zd.Hp, zd.Wp, cout, bn_(cout),
                                               cout, False, zd, CB_OUT_PF)))
                    fwd.append(("bn_fwd", {"bn": p + ".downsample.1", "z": zd, "map": m2, "count": cnt, "eps": 1e-5, "mom": 0.1}))
                    fwd.append(("bn_apply", {"bn": p + ".bn2", "z": z2, "map": m2, "y": out, "relu": 1,
                                             "bn_b": p + ".downsample.1", "z_b": zd}))
                else:
                    fwd.append(("bn_apply", {"bn": p + ".bn2", "z": z2, "map": m2, "y": out, "relu": 1, "res": x}))
                B["x"], B["m1"], B["m2"] = x, m1, m2
                x = out
        fwd.append(("affine", {"n_sc": n_sc}))
        for li in range(len(self.levels)):
            fwd.append(("fuse", {"li": li, "n_sc": n_sc}))
        ch = 0
        H0, W0, _ = self.levels[0]
        for li, dc in enumerate(self.deconvs):
            f = self.lvl[li]["fused"]
            k, cu = dc["k"], dc["cout"]
            zu = dc["zu"]
            fwd.append(conv(self._desc([f, None], dc["f"], self._steps_1x1(dc["cin"]), n_sc, f.Hp, f.Wp, k * k * cu, bn_(cu), cu,
                                       False, zu, CB_OUT_UPSAMPLE, up_k=k)))
            mf = self._map(n_sc, zu, self.cat, y_ch_off=ch)                    # forward: plain PF tensors at full resolution
            fwd.append(("bn_fwd", {"bn": dc["bn"], "z": zu, "map": mf, "count": float(n_sc * H0 * W0), "eps": 1e-3, "mom": 0.01}))
            fwd.append(("bn_apply", {"bn": dc["bn"], "z": zu, "map": mf, "y": self.cat, "relu": 1}))
            # backward: dz is produced in the GEMM's row space (level resolution, columns (a, b, co)); dy / y / z sit at
            # the pixel-shuffled position
            dc["map"] = self._map(n_sc, dc["dzu"], self.cat, c_mod=cu, y_ch_off=ch, up_k=k, z_at_y_pitch=cu)
            dc["ch"] = ch
            ch += cu
        y = self.cat
        sh_in = []
        for s, buf in zip(self.shrink, self.shrink_bufs):
            fwd.append(conv(self._desc([y, None], s["f"], self._steps_3x3_s1(s["cin"], y.Wp), n_sc, y.Hp, y.Wp, s["cout"],
                                       bn_(s["cout"]), s["cout"], True, buf, CB_OUT_PF)))
            sh_in.append(y)
            y = buf
        fwd.append(("head_bias", {}))
        heads = [(t, cn) for t, cn in zip(self.head_out, self.head_cn)]
        fwd.append(conv(self._desc([y, None], self.head_f, self._steps_1x1(self.c_last), n_sc, y.Hp, y.Wp, 32, 32, 32, False,
                                   None, CB_OUT_HEADS, heads=heads)))
        head_in = y
        # ============================================================ backward
        H0, W0, _ = self.levels[0]
        hw0 = self.head_mods[0] + ".weight"
        bwd.append(("zero_grads", {}))
        bwd.append(("heads_pack", {"n_sc": n_sc, "H": H0, "W": W0}))
        # head weights: cls | reg | dir are adjacent in gflat -> the packed [tot][c_last] gradient IS their layout
        bwd += self._wgrad(self.g_pf, [head_in, None], [(0, 0, cb * 64, cb * 64) for cb in range(self.c_last // 64)], n_sc,
                           sum(self.head_cn), "", self.c_last, dst=self.gflat[self.param_off[hw0]:])
        dy = self.d_shrink[-1] if self.shrink else self.d_cat
        bwd.append(conv(self._desc([self.g_pf, None], self.head_d, [(0, 0, 0, 0)], n_sc, dy.Hp, dy.Wp, self.c_last,
                                   bn_(self.c_last), self.c_last, False, dy, CB_OUT_PF)))
        for si in reversed(range(len(self.shrink))):
            s, ybuf, xin = self.shrink[si], self.shrink_bufs[si], sh_in[si]
            dyb, dzb = self.d_shrink[si], self.dz_shrink[si]
            m = self._map(n_sc, dzb, ybuf)
            bwd.append(("bn_bwd", {"bn": s["name"], "has_bn": False, "dy": dyb, "y": ybuf, "relu": 1, "z": None, "map": m,
                                   "count": 1.0, "dz": dzb, "dsum": None, "d_gamma": None, "d_beta": self.G[s["name"] + ".bias"]}))
            Wp = xin.Wp
            boxes = [(0, (t // 3 - 1) * Wp + (t % 3 - 1), cb * 64, t * s["cin"] + cb * 64) for t in range(9)
                     for cb in range(s["cin"] // 64)]
            bwd += self._wgrad(dzb, [xin, None], boxes, n_sc, s["cout"], s["name"] + ".weight", 9 * s["cin"])
            bwd.append(self._unpack(s["name"] + ".weight", "conv", cout=s["cout"], cin=s["cin"], taps=9))
            dx = self.d_shrink[si - 1] if si > 0 else self.d_cat
            nblk = 256 if s["cin"] % 256 == 0 else 128            # the epilogue stages at most 256 bias / channel slots
            for c0 in range(0, s["cin"], nblk):
                bwd.append(conv(self._desc([dzb, None], PackRows(s["d"], c0, nblk), self._steps_dgrad_3x3_s1(s["cout"], dzb.Wp),
                                           n_sc, dx.Hp, dx.Wp, nblk, bn_(nblk), nblk, False, dx, CB_OUT_PF, out_ch_off=c0)))
        for li in reversed(range(len(self.deconvs))):
            dc = self.deconvs[li]
            k, cu, cin = dc["k"], dc["cout"], dc["cin"]
            f, dfu = self.lvl[li]["fused"], self.lvl[li]["d_fused"]
            bwd.append(("bn_bwd", {"bn": dc["bn"], "has_bn": True, "dy": self.d_cat, "y": self.cat, "relu": 1, "z": dc["zu"],
                                   "map": dc["map"], "count": float(n_sc * H0 * W0), "dz": dc["dzu"], "dsum": None, "zmask": True,
                                   "d_gamma": self.G[dc["bn"] + ".weight"], "d_beta": self.G[dc["bn"] + ".bias"]}))
            bwd += self._wgrad(dc["dzu"], [f, None], [(0, 0, cb * 64, cb * 64) for cb in range(cin // 64)], n_sc, k * k * cu,
                               dc["name"], cin)
            bwd.append(self._unpack(dc["name"], "deconv", cin=cin, cu=cu, k=k))
            steps = [(0, cb * 64, cb * 64, 0) for cb in range(k * k * cu // 64)]
            bwd.append(conv(self._desc([dc["dzu"], None], dc["d"], steps, n_sc, dfu.Hp, dfu.Wp, cin, bn_(cin), cin, False, dfu,
                                       CB_OUT_PF)))
        bwd.append(("bucket", {"i": 0}))
        bi = len(self.blocks)
        dx_next = None                                            # gradient of this level's output coming from the level above
        for li in reversed(range(len(self.levels))):
            L = self.lvl[li]
            h, w, c = self.levels[li]
            cnt = float(n_img * h * w)
            bwd.append(("fuse_bwd", {"li": li, "n_sc": n_sc, "n_img": n_img, "addend": dx_next}))
            dcur = L["d_out"]
            nb = self.layer_nums[li]
            for k in reversed(range(nb)):
                bi -= 1
                blk, B = self.blocks[bi], L["blocks"][k]
                p, cin, cout = blk["p"], blk["cin"], blk["cout"]
                xin = B["x"]
                dz, d_o1 = L["dz"], L["d_o1"]
                # out = relu(bn2(z2) + identity): gradient of the sum, through bn2 (and the downsample norm)
                bwd.append(("bn_bwd", {"bn": p + ".bn2", "has_bn": True, "dy": dcur, "y": B["out"], "relu": 1, "z": B["z2"],
                                       "map": B["m2"], "count": cnt, "dz": dz, "dsum": None if blk["has_ds"] else L["dsum"],
                                       "d_gamma": self.G[p + ".bn2.weight"], "d_beta": self.G[p + ".bn2.bias"]}))
                if blk["has_ds"]:
                    bwd.append(("bn_bwd", {"bn": p + ".downsample.1", "has_bn": True, "dy": dcur, "y": B["out"], "relu": 1,
                                           "z": B["zd"], "map": B["m2"], "count": cnt, "dz": L["dzd"], "dsum": None,
                                           "d_gamma": self.G[p + ".downsample.1.weight"],
                                           "d_beta": self.G[p + ".downsample.1.bias"]}))
                Wp = B["o1"].Wp
                boxes = [(0, (t // 3 - 1) * Wp + (t % 3 - 1), cb * 64, t * cout + cb * 64) for t in range(9)
                         for cb in range(cout // 64)]
                bwd += self._wgrad(dz, [B["o1"], None], boxes, n_img, cout, p + ".conv2.weight", 9 * cout)
                bwd.append(self._unpack(p + ".conv2.weight", "conv", cout=cout, cin=cout, taps=9))
                bwd.append(conv(self._desc([dz, None], blk["d2"], self._steps_dgrad_3x3_s1(cout, dz.Wp), n_img, d_o1.Hp, d_o1.Wp,
                                           cout, bn_(cout), cout, False, d_o1, CB_OUT_PF)))
                # o1 = relu(bn1(z1))
                bwd.append(("bn_bwd", {"bn": p + ".bn1", "has_bn": True, "dy": d_o1, "y": B["o1"], "relu": 1, "z": B["z1"],
                                       "map": B["m1"], "count": cnt, "dz": dz, "dsum": None, "zmask": True,
                                       "d_gamma": self.G[p + ".bn1.weight"], "d_beta": self.G[p + ".bn1.bias"]}))
                if blk["stride"] == 2:
                    taps = self._s2_taps()
                    boxes = []
                    for t in range(9):
                        (pr_, dr), (pc_, dc_) = taps[t // 3], taps[t % 3]
                        ro = (pr_ * 2 + pc_) * xin.plane_rows + dr * xin.Wp + dc_
                        boxes += [(0, ro, cb * 64, t * cin + cb * 64) for cb in range(cin // 64)]
                    bwd += self._wgrad(dz, [xin, None], boxes, n_img, cout, p + ".conv1.weight", 9 * cin)
                    bwd.append(self._unpack(p + ".conv1.weight", "conv", cout=cout, cin=cin, taps=9))
                    bwd += self._wgrad(L["dzd"], [xin, None], [(0, 0, cb * 64, cb * 64) for cb in range(cin // 64)], n_img,
                                       cout, p + ".downsample.0.weight", cin)
                    bwd.append(self._unpack(p + ".downsample.0.weight", "conv", cout=cout, cin=cin, taps=1))
                    dX = self.d_canvas if li == 0 else self.lvl[li - 1]["dx_in"]
                    for ph in range(2):
                        for pw in range(2):
                            view = ActView(dX, ph * 2 + pw)
                            steps = self._steps_dgrad_3x3_s2(cout, dz.Wp, ph, pw, cout if (ph == 0 and pw == 0) else 0)
                            a1 = L["dzd"] if (ph == 0 and pw == 0) else None
                            bwd.append(conv(self._desc([dz, a1], blk["d1"], steps, n_img, dz.Hp, dz.Wp, cin, bn_(cin), cin,
                                                       False, view, CB_OUT_PF)))
                    dx_next = dX
                else:
                    Wp = xin.Wp
                    boxes = [(0, (t // 3 - 1) * Wp + (t % 3 - 1), cb * 64, t * cin + cb * 64) for t in range(9)
                             for cb in range(cin // 64)]
                    bwd += self._wgrad(dz, [xin, None], boxes, n_img, cout, p + ".conv1.weight", 9 * cin)
                    bwd.append(self._unpack(p + ".conv1.weight", "conv", cout=cout, cin=cin, taps=9))
                    dnext = L["ga"] if dcur is not L["ga"] else L["gb"]
                    bwd.append(conv(self._desc([dz, None], blk["d1"], self._steps_dgrad_3x3_s1(cout, dz.Wp), n_img, dnext.Hp,
                                               dnext.Wp, cin, bn_(cin), cin, False, dnext, CB_OUT_PF, residual=L["dsum"])))
                    dcur = dnext
            bwd.append(("bucket", {"i": len(self.levels) - li}))
        bwd.append(("pfn_bwd", {"n_img": n_img}))
        bwd.append(("bucket", {"i": len(self.levels) + 1}))
        # the gradient permutations (packed GEMM order -> the parameter's layout in gflat) only have to be done before the
        # bucket's all-reduce: one table-driven launch per bucket instead of one launch per weight tensor
        merged: List[Tuple[str, dict]] = []
        pend: List[dict] = []
        for kind, o in bwd:
            if kind == "permute":
                pend.append(o)
                continue
            if kind == "bucket" and pend:
                merged.append(("permute_batch", {"jobs": pend}))
                pend = []
            merged.append((kind, o))
        assert not pend
        return {"fwd": fwd, "bwd": merged}

    def _alloc_dx_in(self):
        """Input-gradient buffers of the first block of levels 1.. (PS layout of the previous level's output)."""
        for li in range(len(self.levels) - 1):
            h, w, c = self.levels[li]
            self.lvl[li]["dx_in"] = Act(self.max_agents, h, w, c, "ps", self.precise, self.device, self.act_dtype)

    # ------------------------------------------------------------------ execution (GPU)
    def _conv_dispatch(self, o, sp):
        self._launch_ops([("conv", o)], 0, sp)

    def run_ops(self, ops, stream_ptr: Optional[int] = None, upto_bucket=None, on_bucket=None):
        if self.plan_only:
            raise RuntimeError("plan_only engine cannot launch (no CUDA library loaded)")
        lib = self.lib
        sp = stream_ptr if stream_ptr is not None else torch.cuda.current_stream(self.device).cuda_stream
        ck = _lib.check
        for kind, o in ops:
            if kind == "conv":
                self._conv_dispatch(o["desc"], sp)
            elif kind == "wgrad":
                ck(lib.cb_wgrad(C.byref(o["desc"]), 0, sp), "cb_wgrad")
            elif kind == "bn_fwd":
                s = self.bn_slot[o["bn"]]
                z = o["z"]
                # statistics + finalize in one launch (the last CTA closes them): 38 launches less per iteration
                ck(lib.cb_bn_stats_finalize(z.ptr, z.lo_off, C.byref(o["map"]), self.redv(o["bn"], 0).data_ptr(),
                                            self.bn_ticket.data_ptr(), o["count"], o["eps"], o["mom"],
                                            self.P[o["bn"] + ".weight"].data_ptr(), self.P[o["bn"] + ".bias"].data_ptr(),
                                            self.R[o["bn"] + ".running_mean"].data_ptr(),
                                            self.R[o["bn"] + ".running_var"].data_ptr(),
                                            self.bnv(o["bn"], 0).data_ptr(), self.bnv(o["bn"], 1).data_ptr(),
                                            self.bnv(o["bn"], 2).data_ptr(), self.bnv(o["bn"], 3).data_ptr(), sp),
                   "cb_bn_stats_finalize")
            elif kind == "bn_apply":
                z, y = o["z"], o["y"]
                zb, res = o.get("z_b"), o.get("res")
                if res is not None and res.layout != "pf":
                    raise RuntimeError("residual must be PF")
                ck(lib.cb_bn_apply(z.ptr, z.lo_off, self.bnv(o["bn"], 0).data_ptr(), self.bnv(o["bn"], 1).data_ptr(),
                                   zb.ptr if zb is not None else None, zb.lo_off if zb is not None else 0,
                                   self.bnv(o["bn_b"], 0).data_ptr() if zb is not None else None,
                                   self.bnv(o["bn_b"], 1).data_ptr() if zb is not None else None,
                                   res.ptr if res is not None else None, res.C if res is not None else 0,
                                   res.lo_off if res is not None else 0, o["relu"], C.byref(o["map"]), y.ptr, y.lo_off, sp),
                   "cb_bn_apply")
            elif kind == "bn_bwd":
                dy, y, z, dz, dsum = o["dy"], o["y"], o["z"], o["dz"], o["dsum"]
                hb = o["has_bn"]
                sums = self.redv(o["bn"], 1).data_ptr()
                mean = self.bnv(o["bn"], 2).data_ptr() if hb else None
                inv = self.bnv(o["bn"], 3).data_ptr() if hb else None
                zm = bool(o.get("zmask")) and hb               # y = relu(bn(z)) with nothing added: mask from z
                msc = self.bnv(o["bn"], 0).data_ptr() if zm else None
                msh = self.bnv(o["bn"], 1).data_ptr() if zm else None
                ck(lib.cb_bn_bwd_reduce(dy.ptr, dy.lo_off, y.ptr, y.lo_off, o["relu"], z.ptr if hb else None,
                                        z.lo_off if hb else 0, mean, inv, msc, msh, C.byref(o["map"]), sums, sp),
                   "cb_bn_bwd_reduce")
                ck(lib.cb_bn_bwd_apply(dy.ptr, dy.lo_off, y.ptr, y.lo_off, o["relu"], z.ptr if hb else None,
                                       z.lo_off if hb else 0, mean, inv,
                                       self.P[o["bn"] + ".weight"].data_ptr() if hb else None, msc, msh, sums, o["count"],
                                       C.byref(o["map"]), dz.ptr, dz.lo_off, dsum.ptr if dsum is not None else None,
                                       dsum.lo_off if dsum is not None else 0,
                                       o["d_gamma"].data_ptr() if o["d_gamma"] is not None else None,
                                       o["d_beta"].data_ptr() if o["d_beta"] is not None else None, sp), "cb_bn_bwd_apply")
            elif kind == "permute":
                src = o["src"].data_ptr() + o["src_off"] * 4
                R1, R0, K0, st = self.permute_job_shape(o)
                ck(lib.cb_permute_f32(src, R1, R0, 1, K0, st[0], st[1], st[2], st[3], 1.0, o["dst"].data_ptr(), sp),
                   "cb_permute_f32")
            elif kind == "permute_batch":         # all gradient permutations of one all-reduce bucket in one launch
                if "_tab" not in o:
                    arr, total = [], 0
                    for q in o["jobs"]:
                        j = _lib.PackJob()
                        j.src, j.dst = q["src"].data_ptr() + q["src_off"] * 4, q["dst"].data_ptr()
                        R1, R0, K0, st = self.permute_job_shape(q)
                        j.s_r1, j.s_r0, j.s_k1, j.s_k0, j.first = st[0], st[1], st[2], st[3], total
                        j.R0, j.K0, j.rows, j.K = R0, K0, R1 * R0, K0
                        j.dst_ld, j.k_off, j.lo_col_off = K0, 0, -1
                        arr.append(j)
                        total += R1 * R0 * K0
                    buf = (_lib.PackJob * len(arr))(*arr)
                    o["_tab"] = (torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).to(self.device), len(arr), total)
                tab, n_jobs, total = o["_tab"]
                ck(lib.cb_pack_weights_batch(tab.data_ptr(), n_jobs, total, sp), "cb_pack_weights_batch")
            elif kind == "pack_all":
                jobs, n_jobs, total = self._pack_job_table()
                ck(lib.cb_pack_weights_batch(jobs.data_ptr(), n_jobs, total, sp), "cb_pack_weights_batch")
            elif kind == "head_bias":
                c0 = 0
                for h, cn in zip(self.head_mods, self.head_cn):
                    self.head_bias[c0:c0 + cn].copy_(self.P[h + ".bias"])
                    c0 += cn
            elif kind == "pfn_fwd":
                self._pfn_forward(o["n_img"], sp)
            elif kind == "pfn_bwd":
                self._pfn_backward(o["n_img"], sp)
            elif kind == "affine":
                ck(lib.cb_normalize_affine(self.pairwise.data_ptr(), o["n_sc"], self.max_cav, self.ny, self.nx,
                                           float(self.voxel_size[0]), self.affine.data_ptr(), sp), "cb_normalize_affine")
            elif kind == "fuse":
                L = self.lvl[o["li"]]
                src, dst = L["out"], L["fused"]
                h, w, c = self.levels[o["li"]]
                ck(lib.cb_warp_att_fuse(src.ptr, 1 if src.layout == "ps" else 0, src.lo_off, src.n_cap, self.affine.data_ptr(),
                                        self.agent_off.data_ptr(), o["n_sc"], self.max_cav, h, w, c, self.method, dst.ptr,
                                        dst.lo_off, sp), "cb_warp_att_fuse")
            elif kind == "fuse_bwd":
                L = self.lvl[o["li"]]
                src, dfu, dout = L["out"], L["d_fused"], L["d_out"]
                h, w, c = self.levels[o["li"]]
                n_img = o["n_img"]
                L["dfeat"][:n_img * h * w * c].zero_()
                ck(lib.cb_warp_att_fuse_bwd(src.ptr, 1 if src.layout == "ps" else 0, src.lo_off, src.n_cap,
                                            self.affine.data_ptr(), self.agent_off.data_ptr(), o["n_sc"], self.max_cav, h, w, c,
                                            self.method, dfu.ptr, dfu.lo_off, L["dfeat"].data_ptr(), sp), "cb_warp_att_fuse_bwd")
                add = o["addend"]
                ck(lib.cb_grad_combine(L["dfeat"].data_ptr(), add.ptr if add is not None else None,
                                       add.lo_off if add is not None else 0, 1 if dout.layout == "ps" else 0, dout.n_cap, n_img,
                                       h, w, c, dout.ptr, dout.lo_off, sp), "cb_grad_combine")
            elif kind == "zero_grads":
                self.wgflat.zero_()
                self.gflat.zero_()
                self.red_b.zero_()
                self.head_dbias.zero_()
            elif kind == "zero_fwd":
                self.red_f.zero_()
            elif kind == "heads_pack":
                ptrs = (C.c_void_p * 4)(*[t.data_ptr() for t in self.head_grad], *([None] * (4 - len(self.head_grad))))
                cns = (C.c_int32 * 4)(*self.head_cn, *([0] * (4 - len(self.head_cn))))
                ck(lib.cb_heads_grad_pack(ptrs, cns, len(self.head_cn), o["n_sc"], o["H"], o["W"], self.g_pf.ptr, self.g_pf.lo_off,
                                          self.head_dbias.data_ptr(), sp), "cb_heads_grad_pack")
                c0 = 0
                for h, cn in zip(self.head_mods, self.head_cn):
                    self.G[h + ".bias"].copy_(self.head_dbias[c0:c0 + cn])
                    c0 += cn
            elif kind == "bucket":
                if on_bucket is not None:
                    on_bucket(o["i"])
            else:
                raise RuntimeError("unknown op " + kind)

    def _pack_job_table(self):
        """Device table of every weight-packing job (built once: parameter and pack buffers never move)."""
        if getattr(self, "_pack_jobs", None) is None:
            arr, total = [], 0
            for pk in self.packs:
                for (src, R1, R0, K1, K0, s1, s0, k1, k0, ro, ko) in pk.jobs:
                    j = _lib.PackJob()
                    j.src, j.dst = src.data_ptr(), pk.w.data_ptr() + ro * pk.k_total * 2
                    j.s_r1, j.s_r0, j.s_k1, j.s_k0, j.first = s1, s0, k1, k0, total
                    j.R0, j.K0, j.rows, j.K = R0, K0, R1 * R0, K1 * K0
                    j.dst_ld, j.k_off, j.lo_col_off = pk.k_total, ko, (pk.k if self.precise else 0)
                    arr.append(j)
                    total += R1 * R0 * K1 * K0
            buf = (_lib.PackJob * len(arr))(*arr)
            host = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8)
            self._pack_jobs = (host.to(self.device), len(arr), total)
        return self._pack_jobs

    def _pfn_forward(self, n_img: int, sp: int):
        lib, ck = self.lib, _lib.check
        bn = "pillar_vfe.pfn_layers.0.norm"
        cap = self.vox_f.shape[0]
        st = self.red_f[self.pfn_red_off:self.pfn_red_off + 128]
        ck(lib.cb_pfn_train_stats(self.vox_f.data_ptr(), self.vox_c.data_ptr(), self.vox_n.data_ptr(), cap,
                                  self.n_vox_dev.data_ptr(), self.max_pts, self._vsize_f.ctypes.data,
                                  self._center_off_f.ctypes.data, st.data_ptr(), sp), "cb_pfn_train_stats")
        ck(lib.cb_pfn_train_finalize(st.data_ptr(), self.n_vox_dev.data_ptr(), cap, self.max_pts, self.pfn_w.data_ptr(),
                                     self.P[bn + ".weight"].data_ptr(), self.P[bn + ".bias"].data_ptr(), 1e-3, 0.01,
                                     self.R[bn + ".running_mean"].data_ptr(), self.R[bn + ".running_var"].data_ptr(),
                                     self.bnv(bn, 0).data_ptr(), self.bnv(bn, 1).data_ptr(), self.bnv(bn, 2).data_ptr(),
                                     self.bnv(bn, 3).data_ptr(), sp), "cb_pfn_train_finalize")
        self.canvas.zero_()
        ck(lib.cb_pfn_scatter(self.vox_f.data_ptr(), self.vox_c.data_ptr(), self.vox_n.data_ptr(), cap,
                              self.n_vox_dev.data_ptr(), self.max_pts, self.pfn_w.data_ptr(), self.bnv(bn, 0).data_ptr(),
                              self.bnv(bn, 1).data_ptr(), self._vsize_f.ctypes.data, self._center_off_f.ctypes.data, n_img,
                              self.canvas.n_cap, self.ny, self.nx, self.canvas.ptr, self.canvas.lo_off, None, None, sp),
           "cb_pfn_scatter")

    def _pfn_backward(self, n_img: int, sp: int):
        lib, ck = self.lib, _lib.check
        bn = "pillar_vfe.pfn_layers.0.norm"
        cap = self.vox_f.shape[0]
        st = self.red_f[self.pfn_red_off:self.pfn_red_off + 128]
        bs = self.red_b[self.pfn_bsum_off:self.pfn_bsum_off + 768]
        dc = self.d_canvas
        ck(lib.cb_pfn_bwd(self.vox_f.data_ptr(), self.vox_c.data_ptr(), self.vox_n.data_ptr(), cap, self.n_vox_dev.data_ptr(),
                          self.max_pts, self.pfn_w.data_ptr(), self.bnv(bn, 0).data_ptr(), self.bnv(bn, 1).data_ptr(),
                          self.bnv(bn, 2).data_ptr(), self.bnv(bn, 3).data_ptr(), self._vsize_f.ctypes.data,
                          self._center_off_f.ctypes.data, dc.ptr, dc.lo_off, dc.n_cap, self.ny, self.nx, bs.data_ptr(), sp),
           "cb_pfn_bwd")
        ck(lib.cb_pfn_bwd_finalize(st.data_ptr(), bs.data_ptr(), self.n_vox_dev.data_ptr(), cap, self.max_pts,
                                   self.pfn_w.data_ptr(), self.P[bn + ".weight"].data_ptr(), self.bnv(bn, 2).data_ptr(),
                                   self.bnv(bn, 3).data_ptr(), self.G["pillar_vfe.pfn_layers.0.linear.weight"].data_ptr(),
                                   self.G[bn + ".weight"].data_ptr(), self.G[bn + ".bias"].data_ptr(), sp), "cb_pfn_bwd_finalize")

    # ------------------------------------------------------------------ public API
    def set_batch(self, voxel_features, voxel_coords, voxel_num_points, record_len: Sequence[int], pairwise):
        record_len = tuple(int(v) for v in record_len)
        self._set_scene_meta(record_len, pairwise)
        m = int(voxel_features.shape[0])
        if m > self.vox_f.shape[0]:
            raise ValueError(f"{m} voxels exceed the engine capacity max_voxels_total = {self.vox_f.shape[0]}")
        if voxel_features.shape[1] != self.max_pts:
            raise ValueError("voxel_features must be (M, max_pts, 4)")
        self.vox_f[:m].copy_(voxel_features, non_blocking=True)
        self.vox_c[:m].copy_(voxel_coords.to(torch.int32), non_blocking=True)
        self.vox_n[:m].copy_(voxel_num_points.to(torch.int32), non_blocking=True)
        self.n_vox_dev.fill_(m)
        return record_len

    def plan(self, record_len: Tuple[int, ...]):
        key = ("train", record_len)
        ent = self._graphs.get(key)
        if ent is None:
            if "dx_in" not in self.lvl[0] and len(self.levels) > 1:
                self._alloc_dx_in()
            ent = self.build_train_ops(record_len)
            self._graphs[key] = ent
        return ent

    def forward_train(self, voxel_features, voxel_coords, voxel_num_points, record_len, pairwise):
        """Train-mode forward (batch statistics, running stats updated).  Returns the head maps (views of static buffers)."""
        rl = self.set_batch(voxel_features, voxel_coords, voxel_num_points, record_len, pairwise)
        ent = self.plan(rl)
        self.run_ops(ent["fwd"])
        self.num_batches_tracked += 1
        self._last = (rl, ent)
        return {name: t[:len(rl)] for name, t in zip(self.head_names, self.head_out)}

    def backward(self, grads: Dict[str, torch.Tensor], on_bucket=None):
        """grads: d(loss)/d(head maps), fp32 NCHW like the outputs.  Fills gflat (views: self.G[name])."""
        rl, ent = self._last
        n = len(rl)
        for t, name in zip(self.head_grad, self.head_names):
            t[:n].copy_(grads[name])
        self.run_ops(ent["bwd"], on_bucket=on_bucket)
        return self.G

    def state_dict_out(self) -> Dict[str, torch.Tensor]:
        out = {k: v.detach().clone() for k, v in self.P.items()}
        out.update({k: v.detach().clone() for k, v in self.R.items()})
        for b in self.bn_names:
            out[b + ".num_batches_tracked"] = torch.tensor(self.num_batches_tracked, dtype=torch.int64)
        return out
