"""Drop-in nn.Module for the reference's model registry.

`PointPillarCoalignB200` is the state_dict-compatible twin of
/root/reference/opencood/models/point_pillar_baseline_multiscale.py:17-135 (class
PointPillarBaselineMultiscale; `CoAlign` in point_pillar_coalign.py:9-10 is an empty subclass): same
constructor argument (the yaml `model.args` dict), same parameter/buffer names and shapes (so reference
checkpoints load with train_utils.load_saved_model), same `forward(data_dict)` input schema and output
dict.  The forward itself runs entirely on our CUDA library - there is no PyTorch/CPU fallback:
constructing the engine without the built library or without a B200 raises.

`.eval()` forward = CoAlignEngine (eval-mode BatchNorm folded into the packed weights, one CUDA graph per batch
signature).  `.train()` forward = coalign_b200.train_engine.TrainEngine behind a torch.autograd.Function: batch-statistics
BatchNorm with running-stat updates, and `loss.backward()` runs our backward kernels and hands every parameter its gradient,
so the loop body of /root/reference/opencood/tools/train.py:105-125 (zero_grad, forward, criterion, backward,
torch.optim.Adam.step) and DistributedDataParallel's gradient hooks (train_ddp.py:104-109) work unchanged.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn


class _PFN(nn.Module):                         # pillar_vfe.py:8-29 parameter container
    def __init__(self, cin, cout):
        super().__init__()
        self.linear = nn.Linear(cin, cout, bias=False)
        self.norm = nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01)


class _PillarVFE(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        cin = 4 + (6 if cfg["use_absolute_xyz"] else 3) + (1 if cfg["with_distance"] else 0)
        if not cfg["use_norm"] or cfg["with_distance"] or not cfg["use_absolute_xyz"] or len(cfg["num_filters"]) != 1:
            raise NotImplementedError("B200 path implements the CoAlign PillarVFE config "
                                      "(use_norm, use_absolute_xyz, no distance, one PFN layer)")
        self.pfn_layers = nn.ModuleList([_PFN(cin, cfg["num_filters"][0])])


class _BasicBlock(nn.Module):                  # resblock.py:23-51 parameter container
    def __init__(self, cin, cout, stride, has_ds):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        if has_ds:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class _ResNet(nn.Module):                      # resblock.py:135-210
    def __init__(self, layer_nums, strides, filters, inplanes):
        super().__init__()
        for i, (nb, st, pl) in enumerate(zip(layer_nums, strides, filters)):
            blocks = [_BasicBlock(inplanes, pl, st, st != 1 or inplanes != pl)]
            blocks += [_BasicBlock(pl, pl, 1, False) for _ in range(1, nb)]
            setattr(self, f"layer{i}", nn.Sequential(*blocks))
            inplanes = pl
        for m in self.modules():               # resblock.py:165-170
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class _Backbone(nn.Module):                    # base_bev_backbone_resnet.py:15-86
    def __init__(self, cfg):
        super().__init__()
        self.resnet = _ResNet(cfg["layer_nums"], cfg["layer_strides"], cfg["num_filters"], cfg.get("inplanes", 64))
        self.deblocks = nn.ModuleList()
        for cin, cout, s in zip(cfg["num_filters"], cfg["num_upsample_filter"], cfg["upsample_strides"]):
            if s < 1:
                raise NotImplementedError("fractional upsample strides are not used by CoAlign")
            self.deblocks.append(nn.Sequential(nn.ConvTranspose2d(cin, cout, s, stride=s, bias=False),
                                               nn.BatchNorm2d(cout, eps=1e-3, momentum=0.01), nn.ReLU()))


class _DoubleConv(nn.Module):                  # downsample_conv.py:7-27
    def __init__(self, cin, cout, k, s, p):
        super().__init__()
        self.double_conv = nn.Sequential(nn.Conv2d(cin, cout, k, s, p), nn.ReLU(inplace=True),
                                         nn.Conv2d(cout, cout, 3, padding=1), nn.ReLU(inplace=True))


class _Shrink(nn.Module):                      # downsample_conv.py:30-50
    def __init__(self, cfg):
        super().__init__()
        self.layers = nn.ModuleList()
        cin = cfg["input_dim"]
        for k, d, s, p in zip(cfg["kernal_size"], cfg["dim"], cfg["stride"], cfg["padding"]):
            self.layers.append(_DoubleConv(cin, d, k, s, p))
            cin = d


class _TrainStep(torch.autograd.Function):
    """Autograd bridge of the train-mode forward: inputs are the module's parameters (so that autograd, optimizers and DDP
    hooks see them), outputs the head maps; backward runs TrainEngine.backward and returns one gradient per parameter."""

    @staticmethod
    def forward(ctx, module, data_dict, names, *params):
        eng = module._train_engine(data_dict)
        pl = data_dict["processed_lidar"]
        out = eng.forward_train(pl["voxel_features"].float(), pl["voxel_coords"], pl["voxel_num_points"],
                                [int(v) for v in data_dict["record_len"].tolist()], data_dict["pairwise_t_matrix"])
        ctx.eng, ctx.names, ctx.need = eng, names, [p.requires_grad for p in params]
        ctx.head_names = list(eng.head_names)
        for b in module.buffers():                                   # nn.BatchNorm bookkeeping (num_batches_tracked)
            if b.dtype == torch.int64 and b.dim() == 0:
                b += 1
        return tuple(out[k].clone() for k in eng.head_names)

    @staticmethod
    def backward(ctx, *gouts):
        eng = ctx.eng
        grads = {}
        for k, g, ref in zip(ctx.head_names, gouts, eng.head_out):
            grads[k] = g.contiguous().float() if g is not None else torch.zeros_like(ref[:len(eng._last[0])])
        G = eng.backward(grads)
        return (None, None, None) + tuple(G[n].clone() if need else None for n, need in zip(ctx.names, ctx.need))


class PointPillarCoalignB200(nn.Module):
    """core_method: point_pillar_coalign_b200 (registry rule: train_utils.py:127-146)."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        if not args["base_bev_backbone"].get("resnet", True):
            # point_pillar_baseline_multiscale.py:31-34 builds BaseBEVBackbone for resnet: false; CoAlign's yamls all set true
            raise NotImplementedError("point_pillar_coalign_b200 implements base_bev_backbone.resnet: true")
        if "fusion_method" not in args:                    # the reference indexes args['fusion_method'] (:36-41)
            raise KeyError("model.args.fusion_method is required ('att' or 'max')")
        if args["fusion_method"] == "att":
            fd = list(args.get("att", {}).get("feat_dim", args["base_bev_backbone"]["num_filters"]))
            if fd != list(args["base_bev_backbone"]["num_filters"]):
                # ScaledDotProductAttention(feat_dim[i]) scales the scores by 1/sqrt(feat_dim[i]) (att_fuse.py:40-44); the
                # fusion kernel uses the channel count of the level, which every shipped yaml makes equal
                raise NotImplementedError("att.feat_dim must equal base_bev_backbone.num_filters on the B200 path")
        self.pillar_vfe = _PillarVFE(args["pillar_vfe"])
        self.backbone = _Backbone(args["base_bev_backbone"])
        self.fusion_net = nn.ModuleList()            # parameter-free (fusion_in_one.py:91-136)
        out_c = sum(args["base_bev_backbone"]["num_upsample_filter"])
        if "shrink_header" in args:
            self.shrink_conv = _Shrink(args["shrink_header"])
            out_c = args["shrink_header"]["dim"][-1]
        an = args["anchor_number"]
        self.cls_head = nn.Conv2d(out_c, an, 1)
        self.reg_head = nn.Conv2d(out_c, 7 * an, 1)
        if "dir_args" in args:
            self.dir_head = nn.Conv2d(out_c, args["dir_args"]["num_bins"] * an, 1)
        self.max_cav = int(args.get("max_cav", 5))
        self.precise = bool(args.get("b200_precise", False))
        self.block_n_cap = int(args.get("b200_block_n", 256))     # 256-wide tiles run on CTA pairs (the benchmarked engine)
        self._engine = None
        self._engine_key = None
        self._train_eng = None
        if args.get("backbone_fix", False):
            self.backbone_fix()

    def backbone_fix(self):                          # point_pillar_baseline_multiscale.py:68-91
        for name in ("pillar_vfe", "backbone", "shrink_conv", "cls_head", "reg_head"):
            if hasattr(self, name):
                for p in getattr(self, name).parameters():
                    p.requires_grad = False

    # ---- engine management: weights are packed once per (device, weights version, capacity)
    def _weights_version(self):
        return tuple(int(p._version) for p in self.parameters()) + tuple(int(b._version) for b in self.buffers())

    def invalidate(self):
        self._engine = None
        self._train_eng = None

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate()
        return r

    def _apply(self, fn, *a, **k):                   # .to() / .cuda(): parameter storage moves, engines are rebuilt
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def _train_engine(self, data_dict):
        """TrainEngine whose flat parameter / running-stat buffers ARE the storage of this module's parameters and buffers
        (`p.data` is re-pointed once), so optimizer updates reach the kernels without copies."""
        from .train_engine import TrainEngine
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PointPillarCoalignB200 runs on a B200 only: call .to('cuda') (no CPU fallback)")
        rl = [int(v) for v in data_dict["record_len"].tolist()]
        m = int(data_dict["processed_lidar"]["voxel_features"].shape[0])
        e = getattr(self, "_train_eng", None)
        if e is None or e.max_agents < sum(rl) or e.max_scenes < len(rl) or e.max_voxels_total < m or e.device != dev:
            cap_s = max(len(rl), e.max_scenes if e is not None else 1)
            cap_a = max(sum(rl), e.max_agents if e is not None else 1)
            cap_v = max(m * 5 // 4, e.max_voxels_total if e is not None else 1)
            self._train_eng = None
            e = TrainEngine(self.args, self.state_dict(), cap_a, cap_s, device=dev, precise=self.precise,
                            max_cav=self.max_cav, max_voxels_total=cap_v,
                            max_pts=int(data_dict["processed_lidar"]["voxel_features"].shape[1]))
            for name, p in self.named_parameters():
                p.data = e.P[name]
            for name, b in self.named_buffers():
                if name in e.R:
                    b.data = e.R[name]
            self._train_eng = e
        return e

    def engine(self, n_agents: int, n_scenes: int):
        from .engine import CoAlignEngine
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PointPillarCoalignB200 runs on a B200 only: call .to('cuda') (no CPU fallback)")
        key = (dev, self._weights_version())
        e = self._engine
        if e is None or self._engine_key != key or e.max_agents < n_agents or e.max_scenes < n_scenes:
            cap_s = max(n_scenes, e.max_scenes if e is not None else 1)
            cap_a = max(n_agents, e.max_agents if e is not None else 1, cap_s * self.max_cav)
            self._engine = None
            self._engine = CoAlignEngine(self.args, self.state_dict(), cap_a, cap_s, device=dev,
                                         precise=self.precise, max_cav=self.max_cav, block_n_cap=self.block_n_cap)
            self._engine_key = key
        return self._engine

    def forward(self, data_dict: Dict):
        pw = data_dict["pairwise_t_matrix"]
        if pw.shape[1] != self.max_cav:
            self.max_cav = int(pw.shape[1])
            self.invalidate()
        if self.training:
            names, params = zip(*self.named_parameters())
            outs = _TrainStep.apply(self, data_dict, names, *params)
            return dict(zip(self._train_eng.head_names, outs))
        pl = data_dict["processed_lidar"]
        record_len = [int(v) for v in data_dict["record_len"].tolist()]
        eng = self.engine(sum(record_len), len(record_len))
        vc = pl["voxel_coords"]
        vn = pl["voxel_num_points"]
        return eng.forward_voxels(pl["voxel_features"].float(), vc.int() if vc.dtype != torch.int32 else vc,
                                  vn.int() if vn.dtype != torch.int32 else vn, record_len, pw)

    @torch.no_grad()
    def forward_points(self, points, pt_offset, record_len, pairwise_t_matrix, max_pts=32, max_voxels=70000):
        """Extension of the boundary: raw clouds in, voxelisation fused on the GPU."""
        if pairwise_t_matrix.shape[1] != self.max_cav:
            self.max_cav = int(pairwise_t_matrix.shape[1])
            self.invalidate()
        eng = self.engine(sum(record_len), len(record_len))
        return eng.forward_points(points, pt_offset, record_len, pairwise_t_matrix, max_pts, max_voxels)


class _PlainBackbone(nn.Module):               # base_bev_backbone.py:5-94 parameter container (BaseBEVBackbone)
    def __init__(self, cfg, cin=64):
        super().__init__()
        self.blocks = nn.ModuleList()
        self.deblocks = nn.ModuleList()
        c_in = [cin, *cfg["num_filters"][:-1]]
        for i, (n, s, f) in enumerate(zip(cfg["layer_nums"], cfg["layer_strides"], cfg["num_filters"])):
            layers = [nn.ZeroPad2d(1), nn.Conv2d(c_in[i], f, 3, stride=s, padding=0, bias=False),
                      nn.BatchNorm2d(f, eps=1e-3, momentum=0.01), nn.ReLU()]
            for _ in range(n):
                layers += [nn.Conv2d(f, f, 3, padding=1, bias=False), nn.BatchNorm2d(f, eps=1e-3, momentum=0.01), nn.ReLU()]
            self.blocks.append(nn.Sequential(*layers))
        for cin_, cout, s in zip(cfg["num_filters"], cfg["num_upsample_filter"], cfg["upsample_strides"]):
            if s < 1:
                raise NotImplementedError("fractional upsample strides are not on the B200 path")
            self.deblocks.append(nn.Sequential(nn.ConvTranspose2d(cin_, cout, s, stride=s, bias=False),
                                               nn.BatchNorm2d(cout, eps=1e-3, momentum=0.01), nn.ReLU()))
        if len(cfg["upsample_strides"]) > len(cfg["layer_nums"]):
            raise NotImplementedError("extra final deblock is not on the B200 path")


class PointPillarB200(nn.Module):
    """core_method: point_pillar_b200 - state_dict-compatible twin of the single-agent detector
    /root/reference/opencood/models/point_pillar.py:17-84 (BASELINE configs[0]; yaml
    opv2v/lidar_only_with_noise/pointpillar_single.yaml): PillarVFE -> scatter -> BaseBEVBackbone (or the ResNet
    backbone with `base_bev_backbone.resnet: true`) -> shrink header -> cls/reg/dir heads, no fusion: every sample of the
    batch is an independent frame.  Same kernels and engine as the CoAlign model; inference only."""

    def __init__(self, args):
        super().__init__()
        self.args = args
        self.pillar_vfe = _PillarVFE(args["pillar_vfe"])
        self.resnet = bool(args["base_bev_backbone"].get("resnet", False))         # point_pillar.py:27-31
        self.backbone = _Backbone(args["base_bev_backbone"]) if self.resnet else _PlainBackbone(args["base_bev_backbone"])
        out_c = sum(args["base_bev_backbone"]["num_upsample_filter"])
        if "shrink_header" in args:
            self.shrink_conv = _Shrink(args["shrink_header"])
            out_c = args["shrink_header"]["dim"][-1]
        an = args["anchor_number"]
        self.cls_head = nn.Conv2d(out_c, an, 1)
        self.reg_head = nn.Conv2d(out_c, 7 * an, 1)
        if "uncertainty_dim" in args:                                  # point_pillar_uncertainty.py:34-35
            self.unc_head = nn.Conv2d(out_c, int(args["uncertainty_dim"]) * an, 1)
        if "dir_args" in args:
            self.dir_head = nn.Conv2d(out_c, args["dir_args"]["num_bins"] * an, 1)
        self.precise = bool(args.get("b200_precise", False))
        self.block_n_cap = int(args.get("b200_block_n", 256))     # 256-wide tiles run on CTA pairs (the benchmarked engine)
        self._engine = None
        self._engine_key = None

    def _weights_version(self):
        return tuple(int(p._version) for p in self.parameters()) + tuple(int(b._version) for b in self.buffers())

    def invalidate(self):
        self._engine = None

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate()
        return r

    def engine(self, n_frames: int):
        from .engine import CoAlignEngine
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("PointPillarB200 runs on a B200 only: call .to('cuda') (no CPU fallback)")
        key = (dev, self._weights_version())
        e = self._engine
        if e is None or self._engine_key != key or e.max_agents < n_frames:
            cap = max(n_frames, e.max_agents if e is not None else 1)
            self._engine = None
            self._engine = CoAlignEngine(self.args, self.state_dict(), cap, cap, device=dev, precise=self.precise,
                                         block_n_cap=self.block_n_cap, backbone="resnet" if self.resnet else "plain",
                                         fusion=False)
            self._engine_key = key
        return self._engine

    def forward(self, data_dict: Dict):
        if self.training:
            raise NotImplementedError("coalign_b200: training-mode forward/backward is not implemented yet "
                                      "(inference path only; see DESIGN.md)")
        pl = data_dict["processed_lidar"]
        vc, vn = pl["voxel_coords"], pl["voxel_num_points"]
        n = int(vc[:, 0].max().item()) + 1 if vc.shape[0] else 1        # point_pillar_scatter.py:41 (same host sync)
        eng = self.engine(n)
        return eng.forward_voxels(pl["voxel_features"].float(), vc.int() if vc.dtype != torch.int32 else vc,
                                  vn.int() if vn.dtype != torch.int32 else vn, [1] * n, None)

    @torch.no_grad()
    def forward_points(self, points, pt_offset, max_pts=32, max_voxels=70000):
        """Extension of the boundary: raw clouds of n independent frames in, voxelisation fused on the GPU."""
        n = len(pt_offset) - 1
        return self.engine(n).forward_points(points, pt_offset, [1] * n, None, max_pts, max_voxels)


class PointPillarUncertaintyB200(PointPillarB200):
    """core_method: point_pillar_uncertainty_b200 - state_dict-compatible twin of the stage-1 detector
    /root/reference/opencood/models/point_pillar_uncertainty.py:15-76 (yaml
    opv2v/lidar_only_with_noise/coalign/pointpillar_uncertainty.yaml): PillarVFE -> scatter -> BaseBEVBackbone ->
    cls/reg/unc/dir 1x1 heads on the 384-channel decoded map (no shrink header).  Its boxes and `unc_preds`
    (log-variances of x, y, yaw per anchor) are what `pose_graph_pre_calc.py` turns into `stage1_boxes.json` for the
    box-alignment pose graph.  One N=32 head GEMM covers the 2+14+6+4 = 26 output channels."""

    def __init__(self, args):
        if "uncertainty_dim" not in args:
            raise KeyError("point_pillar_uncertainty needs model.args.uncertainty_dim")
        if "shrink_header" in args or args["base_bev_backbone"].get("resnet", False):
            raise NotImplementedError("point_pillar_uncertainty has no shrink header and uses BaseBEVBackbone")
        super().__init__(args)
