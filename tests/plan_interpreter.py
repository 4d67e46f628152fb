"""Reference interpreter of the engine's launch plan (TEST INFRASTRUCTURE, CPU, torch).

Executes the `cb_conv_desc` descriptors exactly as include/coalign_b200.h defines the GEMM
    D[q][n] = sum_steps sum_{kk<64} A_sel[q + row_off][col + kk] * W[n][w_k + kk]   (+ epilogue, output modes)
on the engine's own (CPU-resident) buffers and packed weights, so the host logic - BN folding, weight packing, PF/PS
layouts, K-step tables, residual / fused-downsample wiring, pixel-shuffle and heads epilogues - is checked against the
oracle without a GPU.  The pillar and fusion stages use the oracle's formulas on the engine's layouts.
"""
import numpy as np
import torch

from coalign_b200._lib import CB_OUT_HEADS, CB_OUT_PF, CB_OUT_PS, CB_OUT_UPSAMPLE


def _f32(act):
    """hi (+lo) planes of an Act as float32 [rows, C]."""
    t = act.t.float()
    return t[:act.rows] + (t[act.rows:2 * act.rows] if act.precise else 0)


def _store(act, rows_idx, ch0, vals):
    """Write float32 `vals` [n, c] into Act rows (hi/lo split in precise mode, bf16 rounding otherwise)."""
    hi = vals.to(act.t.dtype)
    c = vals.shape[1]
    act.t[rows_idx, ch0:ch0 + c] = hi
    if act.precise:
        act.t[rows_idx + act.rows, ch0:ch0 + c] = (vals - hi.float()).to(torch.bfloat16)


def nchw_to_act(x, act):
    """dense (n,C,H,W) float32 -> PF / PS rows of `act` (halo untouched)."""
    n, C, H, W = x.shape
    nn, hh, ww = torch.meshgrid(torch.arange(n), torch.arange(H), torch.arange(W), indexing="ij")
    rows = act_rows(act, nn, hh, ww).reshape(-1)
    _store(act, rows, 0, x.permute(0, 2, 3, 1).reshape(-1, C))


def act_rows(act, n, h, w):
    pad = getattr(act, "pad", 1)
    if act.layout == "pf":
        return (n * act.Hp + h + pad) * act.Wp + w + pad
    ph = (h & 1) * 2 + (w & 1)
    return ph * act.plane_rows + (n * act.Hp + (h >> 1) + pad) * act.Wp + (w >> 1) + pad


def act_to_nchw(act, n):
    H, W, C = act.H, act.W, act.C
    nn, hh, ww = torch.meshgrid(torch.arange(n), torch.arange(H), torch.arange(W), indexing="ij")
    rows = act_rows(act, nn, hh, ww).reshape(-1)
    return _f32(act)[rows].reshape(n, H, W, C).permute(0, 3, 1, 2).contiguous()


def run_conv(eng, d):
    reg = eng._by_ptr
    A = [(_f32(reg[d.a_ptr[i]]) if d.a_ptr[i] else None) for i in range(2)]
    pc = reg[d.w_ptr]
    Wt = pc.w.float()                                   # [rows, k_total] (hi | lo)
    rows_total = d.n_img * d.Hp * d.Wp
    q = torch.arange(rows_total)
    acc = torch.zeros(rows_total, d.n_total)
    for i in range(d.n_ksteps):
        st = d.ksteps[i]
        a = A[st.a_sel]
        src = q + st.row_off
        ok = (src >= 0) & (src < a.shape[0]) & (src < d.a_rows[st.a_sel])
        blk = torch.zeros(rows_total, 64)
        blk[ok] = a[src[ok], st.col:st.col + 64]
        acc += blk @ Wt[:d.n_total, st.w_k:st.w_k + 64].t()
    n = q // (d.Hp * d.Wp)
    rem = q % (d.Hp * d.Wp)
    hp, wp = rem // d.Wp, rem % d.Wp
    pad = d.in_pad if d.in_pad else 1
    inter = (hp >= pad) & (hp <= d.Hp - 1 - pad) & (wp >= pad) & (wp <= d.Wp - 1 - pad)
    h, w = hp - pad, wp - pad
    cols = torch.arange(d.n_total)
    bias = reg[d.bias].bias if d.bias in reg else pc.bias          # a K-split launch shares the weights, not the bias
    v = acc + bias[cols % d.cout_mod].view(1, -1)
    if d.residual:
        v = v + _f32(reg[d.residual])[:rows_total, :d.n_total]
    if d.relu:
        v = torch.relu(v)
    qi = q[inter]
    if d.out_mode == CB_OUT_HEADS:
        c0 = 0
        for s in range(d.n_heads):
            t, cn = reg[d.head_out[s]], d.head_cn[s]
            t[n[inter], :, h[inter], w[inter]] = v[inter][:, d.head_c0[s]:d.head_c0[s] + cn]
        return
    out = reg[d.out]
    if d.out_mode == CB_OUT_PF:
        if d.out_Hp == d.Hp and d.out_Wp == d.Wp:
            _store(out, qi, d.out_ch_off, v[inter])
        else:                                               # destination with its own halo: addressed by pixel
            rows = (n * d.out_Hp + h + 1) * d.out_Wp + w + 1
            _store(out, rows[inter], d.out_ch_off, v[inter])
    elif d.out_mode == CB_OUT_PS:
        ph = (h & 1) * 2 + (w & 1)
        rows = ph * d.out_plane_rows + (n * d.out_Hp + (h >> 1) + 1) * d.out_Wp + (w >> 1) + 1
        _store(out, rows[inter], d.out_ch_off, v[inter])
    else:
        k = d.up_k
        for ab in range(k * k):
            a_, b_ = ab // k, ab % k
            rows = (n * d.out_Hp + k * h + a_ + 1) * d.out_Wp + (k * w + b_ + 1)
            _store(out, rows[inter], d.out_ch_off, v[inter][:, ab * d.cout_mod:(ab + 1) * d.cout_mod])


def run_plan(eng, sd, args, batch):
    """Whole forward through the launch plan.  batch: the reference-format dict of torch CPU tensors."""
    from oracle import coalign_oracle as O
    pl = batch["processed_lidar"]
    record_len = [int(v) for v in batch["record_len"]]
    n_img, n_sc = sum(record_len), len(record_len)
    # A3..A5 (oracle formulas) into the engine's canvas layout
    pf = O.pillar_vfe(sd, args, pl["voxel_features"], pl["voxel_coords"], pl["voxel_num_points"])
    canvas = O.scatter(pf, pl["voxel_coords"], n_img, eng.ny, eng.nx)
    eng.canvas.t.zero_()
    full = torch.zeros(eng.canvas.n_cap, 64, eng.ny, eng.nx)
    full[:n_img] = canvas
    nchw_to_act(full, eng.canvas)
    affine = (O.normalize_pairwise_tfm(batch["pairwise_t_matrix"], eng.ny, eng.nx, float(args["voxel_size"][0]))
              if eng.fusion else None)
    for kind, o in eng.build_descs(n_img, n_sc):
        if kind == "conv":
            run_conv(eng, o)
        elif kind == "copy":                                  # no-fusion engines: PS level output -> PF (cb_ps_to_pf)
            src, dst = eng.lvl[o]["out"], eng.lvl[o]["fused"]
            if src.layout == "ps":
                full = torch.zeros(dst.n_cap, dst.C, dst.H, dst.W)
                full[:n_img] = act_to_nchw(src, src.n_cap)[:n_img]
                nchw_to_act(full, dst)
        else:
            src, dst = eng.lvl[o]["out"], eng.lvl[o]["fused"]
            x = act_to_nchw(src, src.n_cap)[:n_img]
            fused = O.att_fusion(x, record_len, affine, args.get("fusion_method", "att"))
            full = torch.zeros(dst.n_cap, dst.C, dst.H, dst.W)
            full[:n_sc] = fused
            nchw_to_act(full, dst)
    return {name: t[:n_sc].clone() for name, t in zip(eng.head_names, eng.head_out)}
