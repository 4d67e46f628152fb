"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the training step of the CoAlign model written out by hand - a
train-mode forward that keeps exactly the tensors a backward pass needs, and a backward made of the explicit pieces the
device kernels of SURVEY 8f row 2 will have to provide (no autograd anywhere in this file):

  * loss gradients w.r.t. the heads                        -> already on the device (cb_pointpillar_loss)
  * dgrad / wgrad / bias reductions of every convolution   (torch.nn.grad.conv2d_input / conv2d_weight = the two GEMMs)
  * ConvTranspose2d(k == s): dgrad is a strided convolution with the same weights, wgrad a pixel-shuffled outer product
  * train-mode BatchNorm: sum(dy), sum(dy * x_hat) reductions, then dx = gamma/sigma * (dy - mean(dy) - x_hat * mean(dy * x_hat))
  * ReLU masks, residual adds
  * attention fusion: soft-max backward over the agents per pixel, d(ego vector) from the scores
  * bilinear warp: scatter-add of the four taps (the adjoint of the gather of torch_transformation_utils.py:322-331)
  * PFN: max-pool argmax routing, BatchNorm1d over (pillar, slot), the 10-feature linear layer

It is pinned by tests/golden/train_small.npz: the gradients of all 127 parameters of the UNMODIFIED reference in `.train()`
mode under the unmodified PointPillarLoss (tests/test_train_oracle_cpu.py).  Reference lines: see oracle/coalign_oracle.py
for the forward of every stage; autograd of those lines is what this file restates.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F
from torch.nn.grad import conv2d_input, conv2d_weight

from . import coalign_oracle as O


# ------------------------------------------------------------------------------------------------ BatchNorm (train)
def bn_fwd(x, gamma, beta, eps, dims):
    shape = [1] * x.dim()
    cdim = [d for d in range(x.dim()) if d not in dims][0]
    shape[cdim] = -1
    mean = x.mean(dims, keepdim=True)
    var = x.var(dims, unbiased=False, keepdim=True)
    inv = 1.0 / torch.sqrt(var + eps)
    xh = (x - mean) * inv
    return xh * gamma.view(shape) + beta.view(shape), (xh, inv, gamma, shape, dims)


def bn_bwd(dy, cache):
    xh, inv, gamma, shape, dims = cache
    n = dy.numel() / gamma.numel()
    dbeta = dy.sum(dims)
    dgamma = (dy * xh).sum(dims)
    dx = gamma.view(shape) * inv * (dy - dbeta.view(shape) / n - xh * dgamma.view(shape) / n)
    return dx, dgamma, dbeta


# ------------------------------------------------------------------------------------------------ conv + BN + ReLU
def conv_bn_fwd(x, w, gamma, beta, eps, stride, pad, relu=True):
    z = F.conv2d(x, w, None, stride, pad)
    y, bc = bn_fwd(z, gamma, beta, eps, (0, 2, 3))
    out = F.relu(y) if relu else y
    return out, (x, w, stride, pad, bc, out if relu else None)


def conv_bn_bwd(dout, cache):
    x, w, stride, pad, bc, out = cache
    dy = dout * (out > 0).to(dout.dtype) if out is not None else dout
    dz, dgamma, dbeta = bn_bwd(dy, bc)
    dw = conv2d_weight(x, w.shape, dz, stride=stride, padding=pad)
    dx = conv2d_input(x.shape, w, dz, stride=stride, padding=pad)
    return dx, dw, dgamma, dbeta


# ------------------------------------------------------------------------------------------------ the whole step
def forward_backward(sd: Dict[str, torch.Tensor], args, data_dict, loss_grad_fn):
    """sd: reference state_dict (plain tensors).  loss_grad_fn(out) -> {'cls_preds','reg_preds','dir_preds'} gradients of
    the loss w.r.t. the head outputs (what cb_pointpillar_loss returns).  Returns (head outputs, {param name: gradient})."""
    g: Dict[str, torch.Tensor] = {}
    pl = data_dict["processed_lidar"]
    vf = pl["voxel_features"].to(sd["pillar_vfe.pfn_layers.0.linear.weight"].dtype)   # float32, or float64 for exactness checks
    vc, vn = pl["voxel_coords"], pl["voxel_num_points"]
    record_len = [int(v) for v in data_dict["record_len"]]
    nx, ny, _nz = [int(v) for v in args["point_pillar_scatter"]["grid_size"]]
    n_agents = sum(record_len)
    bb = args["base_bev_backbone"]

    # ---------------- PFN forward (pillar_vfe.py:105-155, 31-53)
    vx, vy, vz = [float(v) for v in args["voxel_size"]]
    rng = [float(v) for v in args["lidar_range"]]
    cnt = vn.to(vf.dtype).view(-1, 1, 1)
    mean = vf[:, :, :3].sum(1, keepdim=True) / cnt
    cf = vc.to(vf.dtype)
    ctr = torch.stack([cf[:, 3] * vx + (vx / 2 + rng[0]), cf[:, 2] * vy + (vy / 2 + rng[1]), cf[:, 1] * vz + (vz / 2 + rng[2])], 1)
    feats = torch.cat([vf, vf[:, :, :3] - mean, vf[:, :, :3] - ctr.unsqueeze(1)], -1)
    mask = (vn.int().view(-1, 1) > torch.arange(vf.shape[1], dtype=torch.int32).view(1, -1)).unsqueeze(-1).to(vf.dtype)
    feats = feats * mask
    w_pfn = sd["pillar_vfe.pfn_layers.0.linear.weight"]
    lin = feats @ w_pfn.t()                                                          # (M, 32, 64)
    pre = "pillar_vfe.pfn_layers.0.norm"
    y_pfn, bc_pfn = bn_fwd(lin, sd[pre + ".weight"], sd[pre + ".bias"], 1e-3, (0, 1))    # BatchNorm1d over (pillar, slot)
    r_pfn = F.relu(y_pfn)
    pf, arg = r_pfn.max(dim=1)                                                       # (M, 64), slot of the max
    canvas = O.scatter(pf, vc, n_agents, ny, nx)

    # ---------------- encoder forward
    enc_cache: List[list] = []
    feats_lvl = []
    x = canvas
    inpl = bb.get("inplanes", 64)
    for li, (nb, st, planes) in enumerate(zip(bb["layer_nums"], bb["layer_strides"], bb["num_filters"])):
        blocks = []
        for k in range(nb):
            p = f"backbone.resnet.layer{li}.{k}"
            s = st if k == 0 else 1
            has_ds = k == 0 and (s != 1 or inpl != planes)
            o1, c1 = conv_bn_fwd(x, sd[p + ".conv1.weight"], sd[p + ".bn1.weight"], sd[p + ".bn1.bias"], 1e-5, s, 1, True)
            o2, c2 = conv_bn_fwd(o1, sd[p + ".conv2.weight"], sd[p + ".bn2.weight"], sd[p + ".bn2.bias"], 1e-5, 1, 1, False)
            cd = None
            idn = x
            if has_ds:
                idn, cd = conv_bn_fwd(x, sd[p + ".downsample.0.weight"], sd[p + ".downsample.1.weight"],
                                      sd[p + ".downsample.1.bias"], 1e-5, s, 0, False)
            out = F.relu(o2 + idn)
            blocks.append((p, c1, c2, cd, out))
            x = out
        inpl = planes
        enc_cache.append(blocks)
        feats_lvl.append(x)

    # ---------------- fusion forward (warp taps kept for the adjoint)
    affine = O.normalize_pairwise_tfm(data_dict["pairwise_t_matrix"], ny, nx, float(args["voxel_size"][0]))
    method = args.get("fusion_method", "att")
    fused, fuse_cache = [], []
    for f in feats_lvl:
        outs, caches, start = [], [], 0
        for b, n in enumerate(record_len):
            xb = f[start:start + n]
            taps = _warp_taps(xb.shape, affine[b, 0, :n], xb.dtype)
            wv = _warp_apply(xb, taps)
            C = xb.shape[1]
            if method == "max":
                o, am = wv.max(dim=0)
                caches.append((start, n, taps, wv, None, am))
            else:
                score = (wv[0:1] * wv).sum(1) / np.sqrt(C)
                att = torch.softmax(score, dim=0)
                o = (att.unsqueeze(1) * wv).sum(0)
                caches.append((start, n, taps, wv, att, None))
            outs.append(o)
            start += n
        fused.append(torch.stack(outs))
        fuse_cache.append(caches)

    # ---------------- decoder, shrink, heads forward
    ups, dec_cache = [], []
    for i, s in enumerate(bb["upsample_strides"]):
        w = sd[f"backbone.deblocks.{i}.0.weight"]
        z = F.conv_transpose2d(fused[i], w, None, stride=s)
        y, bc = bn_fwd(z, sd[f"backbone.deblocks.{i}.1.weight"], sd[f"backbone.deblocks.{i}.1.bias"], 1e-3, (0, 2, 3))
        u = F.relu(y)
        ups.append(u)
        dec_cache.append((fused[i], w, s, bc, u))
    dec = torch.cat(ups, 1)
    sh_cache = []
    x = dec
    if "shrink_header" in args:
        shc = args["shrink_header"]
        for li, (k, s, p_) in enumerate(zip(shc["kernal_size"], shc["stride"], shc["padding"])):
            pre = f"shrink_conv.layers.{li}.double_conv"
            for idx, (ss, pp) in ((".0", (s, p_)), (".2", (1, 1))):
                y = F.relu(F.conv2d(x, sd[pre + idx + ".weight"], sd[pre + idx + ".bias"], ss, pp))
                sh_cache.append((pre + idx, x, ss, pp, y))
                x = y
    head_in = x
    out = O.heads(sd, head_in)

    # ================================================================== backward
    dout = loss_grad_fn(out)
    dx = torch.zeros_like(head_in)
    for name, key in (("cls_head", "cls_preds"), ("reg_head", "reg_preds"), ("dir_head", "dir_preds")):
        if key not in dout:
            continue
        dy = dout[key]
        w = sd[name + ".weight"]
        g[name + ".bias"] = dy.sum((0, 2, 3))
        g[name + ".weight"] = conv2d_weight(head_in, w.shape, dy)
        dx = dx + conv2d_input(head_in.shape, w, dy)
    for pre, xin, ss, pp, y in reversed(sh_cache):                                    # conv + bias + ReLU
        dz = dx * (y > 0).to(dx.dtype)
        w = sd[pre + ".weight"]
        g[pre + ".bias"] = dz.sum((0, 2, 3))
        g[pre + ".weight"] = conv2d_weight(xin, w.shape, dz, stride=ss, padding=pp)
        dx = conv2d_input(xin.shape, w, dz, stride=ss, padding=pp)
    # decoder: split the concat gradient, ReLU mask, BN, transposed conv
    dfused, c0 = [], 0
    for i, (xin, w, s, bc, u) in enumerate(dec_cache):
        cu = u.shape[1]
        du = dx[:, c0:c0 + cu] * (u > 0).to(dx.dtype)
        c0 += cu
        dz, dgam, dbet = bn_bwd(du, bc)
        g[f"backbone.deblocks.{i}.1.weight"], g[f"backbone.deblocks.{i}.1.bias"] = dgam, dbet
        dfused.append(F.conv2d(dz, w, None, stride=s))                                # adjoint of conv_transpose(k == s)
        # wgrad: dW[ci, co, a, b] = sum_{n,h,w} x[n,ci,h,w] * dz[n,co,h*s+a,w*s+b]
        n_, ci, h, w_ = xin.shape
        dzp = dz.view(n_, cu, h, s, w_, s)
        g[f"backbone.deblocks.{i}.0.weight"] = torch.einsum("nihw,nohawb->ioab", xin, dzp)
    # fusion backward -> gradients of the per-agent level maps
    dfeats = []
    for li, f in enumerate(feats_lvl):
        df = torch.zeros_like(f)
        C = f.shape[1]
        for b, (start, n, taps, wv, att, am) in enumerate(fuse_cache[li]):
            do = dfused[li][b]                                                        # (C, H, W)
            if att is None:                                                           # MaxFusion: route to the arg-max agent
                dwv = torch.zeros_like(wv)
                dwv.scatter_(0, am.unsqueeze(0), do.unsqueeze(0))
            else:
                dwv = att.unsqueeze(1) * do.unsqueeze(0)                              # d(out)/d(w_j) direct term
                datt = (do.unsqueeze(0) * wv).sum(1)                                  # (n, H, W)
                dscore = att * (datt - (att * datt).sum(0, keepdim=True))             # soft-max backward over the agents
                dscore = dscore / np.sqrt(C)
                dwv = dwv + dscore.unsqueeze(1) * wv[0:1]                             # s_j = <w_0, w_j>: d w_j
                dwv[0] = dwv[0] + (dscore.unsqueeze(1) * wv).sum(0)                   #                     d w_0
            df[start:start + n] = _warp_adjoint(dwv, taps)
        dfeats.append(df)
    # encoder backward (deepest level first; a level's input gradient adds to the previous level's map gradient)
    dx = None
    for li in reversed(range(len(enc_cache))):
        dlevel = dfeats[li] if dx is None else dfeats[li] + dx
        dcur = dlevel
        for (p, c1, c2, cd, out_blk) in reversed(enc_cache[li]):
            dsum = dcur * (out_blk > 0).to(dcur.dtype)
            d1, g[p + ".conv2.weight"], g[p + ".bn2.weight"], g[p + ".bn2.bias"] = conv_bn_bwd(dsum, c2)
            dxin, g[p + ".conv1.weight"], g[p + ".bn1.weight"], g[p + ".bn1.bias"] = conv_bn_bwd(d1, c1)
            if cd is not None:
                dds, g[p + ".downsample.0.weight"], g[p + ".downsample.1.weight"], g[p + ".downsample.1.bias"] = conv_bn_bwd(dsum, cd)
                dxin = dxin + dds
            else:
                dxin = dxin + dsum
            dcur = dxin
        dx = dcur
    dcanvas = dx
    # scatter adjoint, max routing, ReLU, BatchNorm1d, linear
    a = vc[:, 0].long()
    idx = (vc[:, 1] + vc[:, 2] * nx + vc[:, 3]).long()
    dpf = dcanvas.view(n_agents, dcanvas.shape[1], ny * nx)[a, :, idx]                # (M, 64)
    dr = torch.zeros_like(r_pfn)
    dr.scatter_(1, arg.unsqueeze(1), dpf.unsqueeze(1))
    dy = dr * (r_pfn > 0).to(dr.dtype)
    dlin, g["pillar_vfe.pfn_layers.0.norm.weight"], g["pillar_vfe.pfn_layers.0.norm.bias"] = bn_bwd(dy, bc_pfn)
    g["pillar_vfe.pfn_layers.0.linear.weight"] = torch.einsum("msc,msf->cf", dlin, feats)
    return out, g


# ------------------------------------------------------------------------------------------------ warp as taps
def _warp_taps(shape, M, dtype):
    """The four bilinear taps of warp_affine_simple (coalign_oracle.warp_affine_simple) as (linear index, weight) pairs."""
    N, _C, H, W = shape
    M = M.double()
    xs = (2.0 * torch.arange(W, dtype=torch.float64) + 1.0) / W - 1.0
    ys = (2.0 * torch.arange(H, dtype=torch.float64) + 1.0) / H - 1.0
    gx = (M[:, 0, 0].view(N, 1, 1) * xs.view(1, 1, W) + M[:, 0, 1].view(N, 1, 1) * ys.view(1, H, 1) + M[:, 0, 2].view(N, 1, 1)).to(dtype)
    gy = (M[:, 1, 0].view(N, 1, 1) * xs.view(1, 1, W) + M[:, 1, 1].view(N, 1, 1) * ys.view(1, H, 1) + M[:, 1, 2].view(N, 1, 1)).to(dtype)
    ix = ((gx + 1) * W - 1) / 2
    iy = ((gy + 1) * H - 1) / 2
    x0, y0 = torch.floor(ix), torch.floor(iy)
    wx1, wy1 = ix - x0, iy - y0
    taps = []
    for dy_, dx_, wgt in ((0, 0, (1 - wy1) * (1 - wx1)), (0, 1, (1 - wy1) * wx1), (1, 0, wy1 * (1 - wx1)), (1, 1, wy1 * wx1)):
        xi, yi = (x0 + dx_).long(), (y0 + dy_).long()
        valid = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        lin = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).view(N, H * W)
        taps.append((lin, (wgt * valid.to(dtype)).view(N, H * W)))
    return taps


def _warp_apply(src, taps):
    N, C, H, W = src.shape
    flat = src.reshape(N, C, H * W)
    out = torch.zeros_like(flat)
    for lin, wgt in taps:
        out = out + torch.gather(flat, 2, lin.unsqueeze(1).expand(N, C, H * W)) * wgt.unsqueeze(1)
    return out.view(N, C, H, W)


def _warp_adjoint(dout, taps):
    """Scatter-add of the output gradient through the four taps (what the backward warp kernel does with atomics)."""
    N, C, H, W = dout.shape
    flat = dout.reshape(N, C, H * W)
    dsrc = torch.zeros_like(flat)
    for lin, wgt in taps:
        dsrc.scatter_add_(2, lin.unsqueeze(1).expand(N, C, H * W), flat * wgt.unsqueeze(1))
    return dsrc.view(N, C, H, W)
