"""BEV half of the camera model (BASELINE configs[4]): the reference's `BevEncodeMSFusion`
(/root/reference/opencood/models/sub_modules/lss_submodule.py:357-417) on the B200 kernels of the LiDAR path.

    splat output (sumN, 128, H, W) -> 7x7/s2 stem + BN + ReLU -> resnet18 layer1..3 (BasicBlocks) -> per-scale AttFusion /
    MaxFusion -> Up(384 -> 256) -> Up(320 -> 256) -> down_layer -> x_fuse (B, 128, H/2, W/2); the same decoder on the un-fused
    per-agent maps -> x_single (sumN, 128, H/2, W/2).

Everything is the existing machinery - tcgen05 implicit-GEMM convolutions (`cb_conv_gemm*`), `cb_warp_att_fuse`, eval-mode
BatchNorm folded into the packed weights - plus three small additions of this round: `cb_conv_desc.in_pad` (the 7x7 stem's
taps reach two pixels into the parity planes of its PS-layout input), `cb_nchw_to_ps_pad` (that input layout) and
`cb_upsample_concat` (`Up`'s bilinear x2 + `torch.cat` written straight into the concat buffer).  Lift + splat
(lift_splat_shoot.py:64-169) is `LiftSplatB200` at the bottom of this file (`cb_lift_splat`); the EfficientNet image encoder
that produces its inputs is not built.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import CB_OUT_PF, CB_OUT_PS
from .engine import BF16, Act, CoAlignEngine, PackedConv, bn_fold, pack_conv_weight


class _ZeroBias:
    """A packed weight matrix with a zero bias (non-final launch of a GEMM that is split over its K-steps)."""

    def __init__(self, pc: PackedConv):
        self.w, self.rows, self.k, self.k_total = pc.w, pc.rows, pc.k, pc.k_total
        self.bias = torch.zeros_like(pc.bias)


class BevEncoderEngine(CoAlignEngine):
    def __init__(self, state_dict: Dict[str, torch.Tensor], H: int, W: int, max_agents: int, max_scenes: int,
                 discrete_ratio: float = 0.4, method: str = "att", device="cuda", precise: bool = False, max_cav: int = 5,
                 block_n_cap: int = 256, plan_only: bool = False, use_graph: bool = True):
        if H % 8 or W % 8:
            raise ValueError("BevEncodeMSFusion on the B200 path needs H, W divisible by 8 (three stride-2 stages)")
        self.plan_only = bool(plan_only)
        self.lib = None if plan_only else _lib.load(check_device=True)
        self.device = torch.device(device)
        self.precise = bool(precise)
        self.max_agents, self.max_scenes, self.max_cav = int(max_agents), int(max_scenes), int(max_cav)
        self.block_n_cap = int(block_n_cap)
        self.use_graph, self.simt_conv, self.pair = bool(use_graph), False, True
        self.chan_major, self.pair_min_bn, self.halo, self.chan_major_256 = True, 256, True, 0
        self.fusion, self.method = True, {"att": 0, "max": 1}[method]
        self.ny, self.nx = H, W
        self.voxel_size = [float(discrete_ratio)] * 3
        self.in_c = int(state_dict["conv1.weight"].shape[1])
        if self.in_c % 64:
            raise NotImplementedError("stem input channels must be a multiple of 64")
        self.levels: List[Tuple[int, int, int]] = [(H // 2, W // 2, 64), (H // 4, W // 4, 128), (H // 8, W // 8, 256)]
        self._pack(state_dict)
        self._alloc_bev()
        self._graphs: "OrderedDict[tuple, dict]" = OrderedDict()
        self.max_graphs = 8
        self._stream = None if plan_only else torch.cuda.Stream(device=self.device)

    # ------------------------------------------------------------------ weights (eval-mode BatchNorm folded)
    def _pack(self, sd):
        sd = {k: v.detach().cpu() for k, v in sd.items()}
        dev, pr = self.device, self.precise

        def conv_bn(wname, bnname):
            s, t = bn_fold(sd, bnname, 1e-5)
            return PackedConv(pack_conv_weight(sd[wname], s), t, pr, dev)

        self.stem = conv_bn("conv1.weight", "bn1")
        self.blocks = []
        cin = 64
        for li, c in ((1, 64), (2, 128), (3, 256)):
            for b in range(2):
                p = f"layer{li}.{b}"
                s1, t1 = bn_fold(sd, p + ".bn1", 1e-5)
                s2, t2 = bn_fold(sd, p + ".bn2", 1e-5)
                w1 = pack_conv_weight(sd[p + ".conv1.weight"], s1)
                w2 = pack_conv_weight(sd[p + ".conv2.weight"], s2)
                has_ds = (p + ".downsample.0.weight") in sd
                b2 = t2
                if has_ds:                                       # 1x1/s2 identity branch: extra K blocks of conv2
                    sdn, tdn = bn_fold(sd, p + ".downsample.1", 1e-5)
                    w2 = torch.cat([w2, pack_conv_weight(sd[p + ".downsample.0.weight"], sdn)], dim=1)
                    b2 = t2 + tdn
                self.blocks.append({"level": li - 1, "k": b, "stride": 2 if (b == 0 and li > 1) else 1,
                                    "cin": cin if b == 0 else c, "cout": c, "has_ds": has_ds,
                                    "c1": PackedConv(w1, t1, pr, dev), "c2": PackedConv(w2, b2, pr, dev)})
            cin = c
        self.ups = {}
        for name in ("up_layer2", "up_layer1"):
            self.ups[name] = (conv_bn(name + ".conv.0.weight", name + ".conv.1"), conv_bn(name + ".conv.3.weight", name + ".conv.4"))
        self.down = [PackedConv(pack_conv_weight(sd[f"down_layer.{i}.weight"], None), sd[f"down_layer.{i}.bias"].double(), pr, dev)
                     for i in (0, 2)]

    # ------------------------------------------------------------------ buffers
    def _alloc_bev(self):
        dev, pr, NA, NS = self.device, self.precise, self.max_agents, self.max_scenes
        A = lambda n, h, w, c, layout="pf", pad=1: Act(n, h, w, c, layout, pr, dev, BF16, pad)      # noqa: E731
        H, W = self.ny, self.nx
        self.x_in = A(NA, H, W, self.in_c, "ps", pad=2)
        (h1, w1, _), (h2, w2, _), (h3, w3, _) = self.levels
        self.stem_tmp = A(NA, h1, w1, 64, "pf", pad=2) if pr else None       # K-split accumulator in the stem's row space
        self.stem_out = A(NA, h1, w1, 64)
        self.lvl = []
        for li, (h, w, c) in enumerate(self.levels):
            last = li == len(self.levels) - 1
            self.lvl.append({"tmp": A(NA, h, w, c), "mid": A(NA, h, w, c), "out": A(NA, h, w, c, "pf" if last else "ps"),
                             "out_pf": None if last else A(NA, h, w, c), "fused": A(NS, h, w, c)})
        n_dec = max(NA, NS)
        self.dec = {"cat2": A(n_dec, h2, w2, 128 + 256), "u2a": A(n_dec, h2, w2, 256), "u2": A(n_dec, h2, w2, 256),
                    "cat1": A(n_dec, h1, w1, 64 + 256), "u1a": A(n_dec, h1, w1, 256), "u1": A(n_dec, h1, w1, 256),
                    "d0": A(n_dec, h1, w1, 256), "x_single": A(NA, h1, w1, 128), "x_fuse": A(NS, h1, w1, 128)}
        self.affine = torch.zeros(NS, self.max_cav, 2, 3, dtype=torch.float64, device=dev)
        self.pairwise = torch.zeros(NS, self.max_cav, self.max_cav, 4, 4, dtype=torch.float64, device=dev)
        self.agent_off = torch.zeros(NS + 1, dtype=torch.int32, device=dev)
        self.in_f32 = torch.zeros(NA, self.in_c, H, W, dtype=torch.float32, device=dev)

    # ------------------------------------------------------------------ launch list
    def _steps_7x7_s2(self, cin: int, src: Act):
        """Tap r of a k7/s2/p3 conv reads input row 2i + r - 3 = parity a, plane row i + d with r - 3 = 2d + a."""
        steps = []
        for r in range(7):
            ar, dr = (r - 3) & 1, (r - 3) >> 1                    # floor division for negatives
            for s in range(7):
                as_, ds = (s - 3) & 1, (s - 3) >> 1
                ro = (ar * 2 + as_) * src.plane_rows + dr * src.Wp + ds
                for cb in range(cin // 64):
                    steps.append((ro, cb * 64, (r * 7 + s) * cin + cb * 64, 0))
        return steps

    def build_bev_ops(self, n_img: int, n_sc: int):
        ops = []
        bn_ = self._bn_for
        conv = lambda d: ops.append(("conv", d))        # noqa: E731
        x = self.x_in
        so = self.stem_out
        steps = self._steps_7x7_s2(self.in_c, x)
        if self.precise:                                  # 3 x 98 K-steps exceed one launch: split, chained through the residual
            half = len(steps) // 2
            d1 = self._desc([x, None], _ZeroBias(self.stem), steps[:half], n_img, x.Hp, x.Wp, 64, 64, 64, False, self.stem_tmp,
                            CB_OUT_PF)
            d1.in_pad = 2
            conv(d1)
            d2 = self._desc([x, None], self.stem, steps[half:], n_img, x.Hp, x.Wp, 64, 64, 64, True, so, CB_OUT_PF,
                            residual=self.stem_tmp)
            d2.in_pad = 2
            conv(d2)
        else:
            d = self._desc([x, None], self.stem, steps, n_img, x.Hp, x.Wp, 64, 64, 64, True, so, CB_OUT_PF)
            d.in_pad = 2
            conv(d)
        x = so
        bi = 0
        for li in range(3):
            L = self.lvl[li]
            for k in range(2):
                blk = self.blocks[bi]
                bi += 1
                cin, cout, st = blk["cin"], blk["cout"], blk["stride"]
                bn = bn_(cout)
                tmp = L["tmp"]
                steps1 = self._steps_3x3_s2(cin, x) if st == 2 else self._steps_3x3_s1(cin, x.Wp)
                conv(self._desc([x, None], blk["c1"], steps1, n_img, tmp.Hp, tmp.Wp, cout, bn, cout, True, tmp, CB_OUT_PF))
                dst = L["out"] if k == 1 else L["mid"]
                steps2 = self._steps_3x3_s1(cout, tmp.Wp)
                a1, res = None, None
                if blk["has_ds"]:
                    steps2 = steps2 + self._steps_1x1(cin, sel=1, k0=9 * cout)
                    a1 = x
                else:
                    res = x
                conv(self._desc([tmp, a1], blk["c2"], steps2, n_img, tmp.Hp, tmp.Wp, cout, bn, cout, True, dst,
                                CB_OUT_PS if dst.layout == "ps" else CB_OUT_PF, residual=res))
                x = dst
        for li in range(3):
            ops.append(("fuse", li))
            if self.lvl[li]["out_pf"] is not None:
                ops.append(("pscopy", {"li": li, "n": n_img}))
        # decoder on the fused maps (n = scenes), then on the per-agent maps (n = agents)
        for which, n in (("fuse", n_sc), ("single", n_img)):
            f = [(L["fused"] if which == "fuse" else (L["out_pf"] if L["out_pf"] is not None else L["out"])) for L in self.lvl]
            D = self.dec
            (h1, w1, _), (h2, w2, _), (h3, w3, _) = self.levels
            ops.append(("ups", {"src": f[1], "n": n, "h": h2, "w": w2, "c": 128, "scale": 1, "dst": D["cat2"], "ch": 0}))
            ops.append(("ups", {"src": f[2], "n": n, "h": h3, "w": w3, "c": 256, "scale": 2, "dst": D["cat2"], "ch": 128}))
            c0, c1 = self.ups["up_layer2"]
            conv(self._desc([D["cat2"], None], c0, self._steps_3x3_s1(384, D["cat2"].Wp), n, D["u2a"].Hp, D["u2a"].Wp, 256,
                            bn_(256), 256, True, D["u2a"], CB_OUT_PF))
            conv(self._desc([D["u2a"], None], c1, self._steps_3x3_s1(256, D["u2a"].Wp), n, D["u2"].Hp, D["u2"].Wp, 256, bn_(256),
                            256, True, D["u2"], CB_OUT_PF))
            ops.append(("ups", {"src": f[0], "n": n, "h": h1, "w": w1, "c": 64, "scale": 1, "dst": D["cat1"], "ch": 0}))
            ops.append(("ups", {"src": D["u2"], "n": n, "h": h2, "w": w2, "c": 256, "scale": 2, "dst": D["cat1"], "ch": 64}))
            c0, c1 = self.ups["up_layer1"]
            conv(self._desc([D["cat1"], None], c0, self._steps_3x3_s1(320, D["cat1"].Wp), n, D["u1a"].Hp, D["u1a"].Wp, 256,
                            bn_(256), 256, True, D["u1a"], CB_OUT_PF))
            conv(self._desc([D["u1a"], None], c1, self._steps_3x3_s1(256, D["u1a"].Wp), n, D["u1"].Hp, D["u1"].Wp, 256, bn_(256),
                            256, True, D["u1"], CB_OUT_PF))
            conv(self._desc([D["u1"], None], self.down[0], self._steps_3x3_s1(256, D["u1"].Wp), n, D["d0"].Hp, D["d0"].Wp, 256,
                            bn_(256), 256, True, D["d0"], CB_OUT_PF))
            out = D["x_fuse"] if which == "fuse" else D["x_single"]
            conv(self._desc([D["d0"], None], self.down[1], self._steps_3x3_s1(256, D["d0"].Wp), n, out.Hp, out.Wp, 128, bn_(128),
                            128, True, out, CB_OUT_PF))
        return ops

    # ------------------------------------------------------------------ execution
    def _launch_bev(self, ops, n_sc: int, sp: int):
        lib, ck = self.lib, _lib.check
        for kind, o in ops:
            if kind in ("conv", "fuse"):
                self._launch_ops([(kind, o)], n_sc, sp)
            elif kind == "pscopy":
                L = self.lvl[o["li"]]
                h, w, c = self.levels[o["li"]]
                ck(lib.cb_ps_to_pf(L["out"].ptr, L["out"].lo_off, L["out"].n_cap, o["n"], h, w, c, L["out_pf"].ptr,
                                   L["out_pf"].lo_off, sp), "cb_ps_to_pf")
            elif kind == "ups":
                s, d = o["src"], o["dst"]
                ck(lib.cb_upsample_concat(s.ptr, s.lo_off, o["n"], o["h"], o["w"], o["c"], o["scale"], d.ptr, d.lo_off, d.C,
                                          o["ch"], sp), "cb_upsample_concat")
            else:
                raise RuntimeError("unknown op " + kind)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, record_len: Sequence[int], pairwise: torch.Tensor, clone: bool = True,
                channels_last: bool = False):
        """x (sumN, inC, H, W) float32 on the device (the splat output; channels_last: (sumN, H, W, inC), what
        LiftSplatB200 produces); returns (x_single, x_fuse) float32 NCHW."""
        record_len = tuple(int(v) for v in record_len)
        n_img, n_sc = sum(record_len), len(record_len)
        want = (n_img, self.ny, self.nx, self.in_c) if channels_last else (n_img, self.in_c, self.ny, self.nx)
        if tuple(x.shape) != want or x.dtype != torch.float32:
            raise ValueError("x must be float32 %s" % (want,))
        self._set_scene_meta(record_len, pairwise)
        self.in_f32.view(-1)[:x.numel()].copy_(x.reshape(-1), non_blocking=True)
        key = ("bev", record_len, bool(channels_last))
        ent = self._graphs.get(key)
        if ent is None:
            ent = {"ops": self.build_bev_ops(n_img, n_sc), "graph": None}
            self._graphs[key] = ent
            while len(self._graphs) > self.max_graphs:
                self._graphs.popitem(last=False)

        def run(sp):
            to_ps = self.lib.cb_nhwc_to_ps_pad if channels_last else self.lib.cb_nchw_to_ps_pad
            _lib.check(to_ps(self.in_f32.data_ptr(), n_img, self.in_c, self.ny, self.nx, 2, self.x_in.n_cap, self.x_in.ptr,
                             self.x_in.lo_off, sp), "cb_n*_to_ps_pad")
            _lib.check(self.lib.cb_normalize_affine(self.pairwise.data_ptr(), n_sc, self.max_cav, self.ny, self.nx,
                                                    float(self.voxel_size[0]), self.affine.data_ptr(), sp), "cb_normalize_affine")
            self._launch_bev(ent["ops"], n_sc, sp)

        cur = torch.cuda.current_stream(self.device)
        if not self.use_graph:
            run(cur.cuda_stream)
        else:
            if ent["graph"] is None:
                run(cur.cuda_stream)
                cur.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._stream):
                    run(torch.cuda.current_stream(self.device).cuda_stream)
                ent["graph"] = g
            ent["graph"].replay()
        xs = self.read_act(self.dec["x_single"], n_img)
        xf = self.read_act(self.dec["x_fuse"], n_sc)
        return xs, xf


# ----------------------------------------------------------------------------------------------------------------------
# drop-in twin of the reference sub-module
# ----------------------------------------------------------------------------------------------------------------------
import torch.nn as nn       # noqa: E402


class _Block(nn.Module):                          # torchvision BasicBlock parameter container
    def __init__(self, cin, cout, has_ds):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, 2 if has_ds else 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        if has_ds:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, 2, bias=False), nn.BatchNorm2d(cout))


class _Up(nn.Module):                             # lss_submodule.Up parameter container (conv.0/1/3/4)
    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                                  nn.Conv2d(cout, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class BevEncodeMSFusionB200(nn.Module):
    """state_dict-compatible twin of the reference's `BevEncodeMSFusion` (lss_submodule.py:357-417): same constructor
    argument (`fusion_args` with `core_method` in {att_ms, max_ms} and `args.in_channels` / `args.voxel_size`), same 118
    state_dict entries, same `forward(x, record_len, pairwise_t_matrix) -> (x_single, x_fuse)`.  Inference only, B200 only."""

    def __init__(self, fusion_args):
        super().__init__()
        args = fusion_args["args"]
        self.method = {"att_ms": "att", "max_ms": "max"}.get(fusion_args["core_method"])
        if self.method is None:
            raise NotImplementedError("BevEncodeMSFusion: core_method must be att_ms or max_ms")     # the reference raises too
        self.discrete_ratio = float(args["voxel_size"][0])
        self.precise = bool(args.get("b200_precise", False))
        inC = int(args["in_channels"])
        self.conv1 = nn.Conv2d(inC, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.layer1 = nn.Sequential(_Block(64, 64, False), _Block(64, 64, False))
        self.layer2 = nn.Sequential(_Block(64, 128, True), _Block(128, 128, False))
        self.layer3 = nn.Sequential(_Block(128, 256, True), _Block(256, 256, False))
        self.up_layer1 = _Up(64 + 256, 256)
        self.up_layer2 = _Up(128 + 256, 256)
        self.down_layer = nn.Sequential(nn.Conv2d(256, 256, 3, 1, 1), nn.ReLU(inplace=True), nn.Conv2d(256, 128, 3, 1, 1),
                                        nn.ReLU(inplace=True))
        self._eng, self._key = None, None

    def forward(self, x, record_len, pairwise_t_matrix):
        if self.training:
            raise NotImplementedError("BevEncodeMSFusionB200 is inference-only (the training kernels cover the LiDAR model)")
        rl = [int(v) for v in (record_len.tolist() if torch.is_tensor(record_len) else record_len)]
        _, _c, H, W = x.shape
        key = (x.device, H, W, int(pairwise_t_matrix.shape[1]), tuple(int(p._version) for p in self.parameters()))
        e = self._eng
        if e is None or self._key != key or e.max_agents < sum(rl) or e.max_scenes < len(rl):
            self._eng = None
            e = BevEncoderEngine(self.state_dict(), H, W, max(sum(rl), e.max_agents if e is not None else 1),
                                 max(len(rl), e.max_scenes if e is not None else 1), discrete_ratio=self.discrete_ratio,
                                 method=self.method, device=x.device, precise=self.precise,
                                 max_cav=int(pairwise_t_matrix.shape[1]))
            self._eng, self._key = e, key
        return e.forward(x.float().contiguous(), rl, pairwise_t_matrix)


# ----------------------------------------------------------------------------------------------------------------------
# lift + splat (the camera front-end between the image encoder and the BEV encoder)
# ----------------------------------------------------------------------------------------------------------------------
class LiftSplatB200:
    """Device replacement of `LiftSplatShoot.get_geometry` + the lift of `CamEncode.forward` + `voxel_pooling`
    (/root/reference/opencood/models/lift_splat_shoot.py:64-169, lss_submodule.py:134-136): depth logits and image features of
    every camera in, BEV feature map out.  Constructor arguments are the yaml's `grid_conf`, `data_aug_conf.final_dim` and
    `img_downsample` (lss_coalign_fusion.yaml:28-43,99).  The geometry's 3x3 inverses are host plumbing (torch.inverse on B*N
    tiny matrices, as in the reference); everything per frustum point runs in `cb_lift_splat`."""

    def __init__(self, grid_conf: dict, final_dim, downsample: int, device="cuda"):
        self.lib = _lib.load(check_device=True)
        self.device = torch.device(device)
        gc = grid_conf
        self.dx = np.asarray([row[2] for row in (gc["xbound"], gc["ybound"], gc["zbound"])], np.float32)          # gen_dx_bx
        self.bx = np.asarray([row[0] + row[2] / 2.0 for row in (gc["xbound"], gc["ybound"], gc["zbound"])], np.float32)
        self.nx = np.asarray([int((row[1] - row[0]) / row[2]) for row in (gc["xbound"], gc["ybound"], gc["zbound"])], np.int32)
        ogfH, ogfW = final_dim
        self.fH, self.fW = ogfH // downsample, ogfW // downsample
        d_min, d_max, nb = gc["ddiscr"]
        if gc["mode"] == "UD":
            ds = d_min + (d_max - d_min) / nb * np.arange(nb)
        elif gc["mode"] == "LID":
            ds = d_min + 2 * (d_max - d_min) / (nb * (1 + nb)) * (np.arange(nb) * np.arange(1, 1 + nb)) / 2
        else:
            raise NotImplementedError(gc["mode"])
        self.D = int(nb)
        self.ds = torch.tensor(ds, dtype=torch.float).to(self.device)                                   # create_frustum (:69-74)
        self.xs = torch.linspace(0, ogfW - 1, self.fW, dtype=torch.float).to(self.device)
        self.ys = torch.linspace(0, ogfH - 1, self.fH, dtype=torch.float).to(self.device)

    @torch.no_grad()
    def __call__(self, depth_logit, x_img, rots, trans, intrins, post_rots, post_trans, out: torch.Tensor = None):
        """depth_logit (B*N, D, fH, fW), x_img (B*N, C, fH, fW) float32 CUDA; camera tensors (B, N, 3[, 3]).  Returns the BEV
        accumulator (B, ny, nx, nz*C) float32 channels-last (== voxel_pooling's output permuted to NHWC)."""
        B, N = int(trans.shape[0]), int(trans.shape[1])
        BN, C = int(x_img.shape[0]), int(x_img.shape[1])
        if BN != B * N or tuple(depth_logit.shape) != (BN, self.D, self.fH, self.fW) or tuple(x_img.shape[2:]) != (self.fH, self.fW):
            raise ValueError("lift_splat: inconsistent shapes")
        f = lambda t: t.to(self.device, torch.float32)      # noqa: E731
        cam = torch.cat([torch.inverse(f(post_rots)).reshape(BN, 9), f(post_trans).reshape(BN, 3),
                         f(rots).matmul(torch.inverse(f(intrins))).reshape(BN, 9), f(trans).reshape(BN, 3)], 1).contiguous()
        ny, nxx, nz = int(self.nx[1]), int(self.nx[0]), int(self.nx[2])
        if out is None:
            out = torch.zeros(B, ny, nxx, nz * C, dtype=torch.float32, device=self.device)
        else:
            out.zero_()
        dl, xi = depth_logit.float().contiguous(), x_img.float().contiguous()
        _lib.check(self.lib.cb_lift_splat(dl.data_ptr(), xi.data_ptr(), cam.data_ptr(), self.xs.data_ptr(), self.ys.data_ptr(),
                                          self.ds.data_ptr(), B, N, self.D, self.fH, self.fW, C, self.dx.ctypes.data,
                                          self.bx.ctypes.data, self.nx.ctypes.data, out.data_ptr(),
                                          torch.cuda.current_stream(self.device).cuda_stream), "cb_lift_splat")
        return out
