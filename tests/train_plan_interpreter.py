"""CPU interpreter of the training launch plan (TEST INFRASTRUCTURE, torch on the host).

Executes the (kind, dict) op lists of coalign_b200.train_engine.TrainEngine.build_train_ops exactly as include/coalign_b200.h
defines each kernel - conv GEMM descriptors (forward and input-gradient launches), cb_wgrad unit / box lists, the cb_map row
mappings of the BatchNorm kernels, weight packing jobs and gradient permutations - on the engine's own CPU-resident buffers.
What it checks is host logic: layouts, K-step tables with negated shifts, parity-plane dgrad launches, packed-weight
orderings, buffer wiring.  The fusion and PFN stages use the oracle's formulas on the engine's layouts.
"""
import numpy as np
import torch

from coalign_b200._lib import CB_OUT_PF, CB_OUT_PS, CB_OUT_UPSAMPLE
from oracle import backward_oracle as BO
from oracle import coalign_oracle as O
from tests import plan_interpreter as PI


# ------------------------------------------------------------------------------------------------ Act access (views too)
def f32(act):
    """float32 [rows, C] of an Act or of a parity-plane view, hi + lo."""
    if hasattr(act, "row0"):
        p = act.parent
        t = p.t.float()
        v = t[act.row0:p.rows]
        return v + (t[p.rows + act.row0:2 * p.rows] if p.precise else 0)
    return PI._f32(act)


def store(act, rows_idx, ch0, vals):
    if hasattr(act, "row0"):
        PI._store(act.parent, rows_idx + act.row0, ch0, vals)
    else:
        PI._store(act, rows_idx, ch0, vals)


def run_conv(eng, d):
    """PI.run_conv with view-aware operand access."""
    reg = eng._by_ptr
    orig_f32, orig_store = PI._f32, PI._store
    PI._f32 = lambda a: f32(a) if hasattr(a, "row0") else orig_f32(a)
    PI._store = lambda a, r, c, v: store(a, r, c, v) if hasattr(a, "row0") else orig_store(a, r, c, v)
    try:
        PI.run_conv(eng, d)
    finally:
        PI._f32, PI._store = orig_f32, orig_store


# ------------------------------------------------------------------------------------------------ cb_map
def map_index(m):
    """For every interior z row q and column col: (q [Nq], y_row [Nq, c_total], y_ch [c_total])."""
    rows_total = m.n_img * m.Hp * m.Wp
    q = torch.arange(rows_total)
    n = q // (m.Hp * m.Wp)
    rem = q % (m.Hp * m.Wp)
    hp, wp = rem // m.Wp, rem % m.Wp
    inter = (hp >= 1) & (hp <= m.Hp - 2) & (wp >= 1) & (wp <= m.Wp - 2)
    q, n, h, w = q[inter], n[inter], hp[inter] - 1, wp[inter] - 1
    col = torch.arange(m.c_total)
    if m.y_mode == CB_OUT_PF:
        yrow = q.view(-1, 1).expand(-1, m.c_total)
        ych = m.y_ch_off + col
    elif m.y_mode == CB_OUT_PS:
        ph = (h & 1) * 2 + (w & 1)
        r = ph * m.y_plane_rows + (n * m.y_Hp + (h >> 1) + 1) * m.y_Wp + (w >> 1) + 1
        yrow = r.view(-1, 1).expand(-1, m.c_total)
        ych = m.y_ch_off + col
    else:
        ab = col // m.c_mod
        a, b = ab // m.up_k, ab % m.up_k
        yrow = (n.view(-1, 1) * m.y_Hp + m.up_k * h.view(-1, 1) + a.view(1, -1) + 1) * m.y_Wp + m.up_k * w.view(-1, 1) + b.view(1, -1) + 1
        ych = m.y_ch_off + col % m.c_mod
    return q, yrow, ych.view(1, -1).expand(q.shape[0], -1)


def gather_y(act, yrow, ych):
    return f32(act)[yrow, ych]


def scatter_y(act, yrow, ych, vals):
    """write float32 vals [Nq, c_total] into act at (yrow, ych) with the bf16 (+lo) split of the store path."""
    hi = vals.to(act.t.dtype)
    act.t[yrow, ych] = hi
    if act.precise:
        act.t[yrow + act.rows, ych] = (vals - hi.float()).to(torch.bfloat16)


def z_values(m, z, q, yrow, ych):
    if m.z_at_y:
        return f32(z)[yrow, ych - m.y_ch_off]
    return f32(z)[q]


# ------------------------------------------------------------------------------------------------ ops
def emu_pack_all(eng):
    for pk in eng.packs:
        for (src, R1, R0, K1, K0, s1, s0, k1, k0, ro, ko) in pk.jobs:
            flat = src.reshape(-1)
            r1, r0, kk1, kk0 = torch.meshgrid(torch.arange(R1), torch.arange(R0), torch.arange(K1), torch.arange(K0), indexing="ij")
            v = flat[(r1 * s1 + r0 * s0 + kk1 * k1 + kk0 * k0).reshape(-1)].reshape(R1 * R0, K1 * K0)
            hi = v.to(pk.w.dtype)
            pk.w[ro:ro + R1 * R0, ko:ko + K1 * K0] = hi
            if eng.precise:
                pk.w[ro:ro + R1 * R0, pk.k + ko:pk.k + ko + K1 * K0] = (v - hi.float()).to(torch.bfloat16)


def emu_bn_fwd(eng, o):
    m, bn = o["map"], o["bn"]
    q, yrow, ych = map_index(m)
    z = z_values(m, o["z"], q, yrow, ych).double()                        # forward maps never use z_at_y
    c = m.c_mod
    zc = z.reshape(z.shape[0], m.c_total // c, c)
    cnt = zc.shape[0] * zc.shape[1]
    assert abs(cnt - o["count"]) < 0.5, (bn, cnt, o["count"])
    mean = zc.mean((0, 1))
    var = zc.var((0, 1), unbiased=False)
    inv = 1.0 / torch.sqrt(var + o["eps"])
    sc = eng.P[bn + ".weight"].double() * inv
    eng.bnv(bn, 0).copy_(sc.float())
    eng.bnv(bn, 1).copy_((eng.P[bn + ".bias"].double() - mean * sc).float())
    eng.bnv(bn, 2).copy_(mean.float())
    eng.bnv(bn, 3).copy_(inv.float())
    mom = o["mom"]
    rm, rv = eng.R[bn + ".running_mean"], eng.R[bn + ".running_var"]
    rm.copy_(((1 - mom) * rm.double() + mom * mean).float())
    rv.copy_(((1 - mom) * rv.double() + mom * var * cnt / (cnt - 1)).float())


def emu_bn_apply(eng, o):
    m, bn = o["map"], o["bn"]
    q, yrow, ych = map_index(m)
    rep = m.c_total // m.c_mod
    v = f32(o["z"])[q] * eng.bnv(bn, 0).repeat(rep) + eng.bnv(bn, 1).repeat(rep)
    if o.get("z_b") is not None:
        v = v + f32(o["z_b"])[q] * eng.bnv(o["bn_b"], 0).repeat(rep) + eng.bnv(o["bn_b"], 1).repeat(rep)
    if o.get("res") is not None:
        v = v + f32(o["res"])[q][:, :m.c_total]
    if o["relu"]:
        v = torch.relu(v)
    scatter_y(o["y"], yrow, ych, v)


def emu_bn_bwd(eng, o):
    m, bn = o["map"], o["bn"]
    q, yrow, ych = map_index(m)
    g = gather_y(o["dy"], yrow, ych)
    if o["relu"]:
        g = g * (gather_y(o["y"], yrow, ych) > 0).float()
    c, rep = m.c_mod, m.c_total // m.c_mod
    if o["dsum"] is not None:
        store(o["dsum"], q, 0, g)
    if o["has_bn"]:
        mean, inv = eng.bnv(bn, 2).repeat(rep), eng.bnv(bn, 3).repeat(rep)
        xh = (z_values(m, o["z"], q, yrow, ych) - mean) * inv
        db = g.double().reshape(-1, rep, c).sum((0, 1))
        dg = (g.double() * xh.double()).reshape(-1, rep, c).sum((0, 1))
        cnt = o["count"]
        dz = (eng.P[bn + ".weight"].repeat(rep) * inv) * (g - (db / cnt).float().repeat(rep) - xh * (dg / cnt).float().repeat(rep))
        if o["d_gamma"] is not None:
            o["d_gamma"].copy_(dg.float())
    else:
        db = g.double().reshape(-1, rep, c).sum((0, 1))
        dz = g
    if o["d_beta"] is not None:
        o["d_beta"].copy_(db.float())
    store(o["dz"], q, 0, dz)


def emu_wgrad(eng, o):
    dz = f32(o["dz"])
    rows_total = o["desc"].rows_total
    xs = [f32(x) if x is not None else None for x in o["xs"]]
    dst = o["dst"]
    q = torch.arange(rows_total)
    for (m0, mv, boxes) in o["units"]:
        a = torch.zeros(rows_total, 128)
        cols = min(128, dz.shape[1] - m0)
        a[:, :cols] = dz[:rows_total, m0:m0 + cols]
        for (sel, ro, col, oc) in boxes:
            x = xs[sel]
            src = q + ro
            ok = (src >= 0) & (src < x.shape[0])
            b = torch.zeros(rows_total, 64)
            b[ok] = x[src[ok], col:col + 64]
            d = (a.double().t() @ b.double()).float()                     # [128][64]
            base = o["dst_off"] + m0 * o["ld"] + oc
            for mm in range(mv):
                dst[base + mm * o["ld"]: base + mm * o["ld"] + 64] += d[mm]


def emu_permute(eng, o):
    src = o["src"]
    off = o["src_off"]
    if o["kind"] == "conv":
        co, ci, t = o["cout"], o["cin"], o["taps"]
        packed = src[off:off + co * t * ci].view(co, t, ci)
        o["dst"].copy_(packed.permute(0, 2, 1).reshape(o["dst"].shape))
    else:
        ci, cu, k = o["cin"], o["cu"], o["k"]
        packed = src[off:off + k * k * cu * ci].view(k * k, cu, ci)       # [(ab)][co][ci]
        o["dst"].copy_(packed.permute(2, 1, 0).reshape(o["dst"].shape))


def emu_fuse(eng, o, record_len, affine, method):
    L = eng.lvl[o["li"]]
    src, dst = L["out"], L["fused"]
    n_img = sum(record_len)
    x = PI.act_to_nchw(src, src.n_cap)[:n_img]
    fused = O.att_fusion(x, record_len, affine, method)
    full = torch.zeros(dst.n_cap, dst.C, dst.H, dst.W)
    full[:len(record_len)] = fused
    PI.nchw_to_act(full, dst)


def emu_fuse_bwd(eng, o, record_len, affine, method):
    L = eng.lvl[o["li"]]
    src, dfu, dout = L["out"], L["d_fused"], L["d_out"]
    n_img = sum(record_len)
    f = PI.act_to_nchw(src, src.n_cap)[:n_img]
    do_all = PI.act_to_nchw(dfu, dfu.n_cap)[:len(record_len)]
    df = torch.zeros_like(f)
    C = f.shape[1]
    start = 0
    for b, n in enumerate(record_len):
        xb = f[start:start + n]
        taps = BO._warp_taps(xb.shape, affine[b, 0, :n], xb.dtype)
        wv = BO._warp_apply(xb, taps)
        do = do_all[b]
        if method == "max":
            _, am = wv.max(dim=0)
            dwv = torch.zeros_like(wv)
            dwv.scatter_(0, am.unsqueeze(0), do.unsqueeze(0))
        else:
            score = (wv[0:1] * wv).sum(1) / np.sqrt(C)
            att = torch.softmax(score, dim=0)
            dwv = att.unsqueeze(1) * do.unsqueeze(0)
            datt = (do.unsqueeze(0) * wv).sum(1)
            dscore = att * (datt - (att * datt).sum(0, keepdim=True)) / np.sqrt(C)
            dwv = dwv + dscore.unsqueeze(1) * wv[0:1]
            dwv[0] = dwv[0] + (dscore.unsqueeze(1) * wv).sum(0)
        df[start:start + n] = BO._warp_adjoint(dwv, taps)
        start += n
    if o["addend"] is not None:
        df = df + PI.act_to_nchw(o["addend"], o["addend"].n_cap)[:n_img]
    full = torch.zeros(dout.n_cap, dout.C, dout.H, dout.W)
    full[:n_img] = df
    PI.nchw_to_act(full, dout)


def _pfn_feats(eng, args, vf, vc, vn):
    vx, vy, vz = [float(v) for v in args["voxel_size"]]
    rng = [float(v) for v in args["lidar_range"]]
    cnt = vn.to(vf.dtype).view(-1, 1, 1)
    mean = vf[:, :, :3].sum(1, keepdim=True) / cnt
    cf = vc.to(vf.dtype)
    ctr = torch.stack([cf[:, 3] * vx + (vx / 2 + rng[0]), cf[:, 2] * vy + (vy / 2 + rng[1]), cf[:, 1] * vz + (vz / 2 + rng[2])], 1)
    feats = torch.cat([vf, vf[:, :, :3] - mean, vf[:, :, :3] - ctr.unsqueeze(1)], -1)
    mask = (vn.int().view(-1, 1) > torch.arange(vf.shape[1], dtype=torch.int32).view(1, -1)).unsqueeze(-1).to(vf.dtype)
    return feats * mask


def emu_pfn_fwd(eng, args, batch, n_img, cache):
    pl = batch["processed_lidar"]
    vf, vc, vn = pl["voxel_features"].float(), pl["voxel_coords"], pl["voxel_num_points"]
    bn = "pillar_vfe.pfn_layers.0.norm"
    feats = _pfn_feats(eng, args, vf, vc, vn)
    lin = feats @ eng.P["pillar_vfe.pfn_layers.0.linear.weight"].t()
    y, bc = BO.bn_fwd(lin, eng.P[bn + ".weight"], eng.P[bn + ".bias"], 1e-3, (0, 1))
    with torch.no_grad():
        mean, var = lin.mean((0, 1)), lin.var((0, 1), unbiased=True)
        eng.R[bn + ".running_mean"].mul_(0.99).add_(0.01 * mean)
        eng.R[bn + ".running_var"].mul_(0.99).add_(0.01 * var)
    r = torch.relu(y)
    pf, arg = r.max(dim=1)
    cache.update(feats=feats, bc=bc, r=r, arg=arg, vc=vc)
    canvas = O.scatter(pf, vc, n_img, eng.ny, eng.nx)
    eng.canvas.t.zero_()
    full = torch.zeros(eng.canvas.n_cap, 64, eng.ny, eng.nx)
    full[:n_img] = canvas
    PI.nchw_to_act(full, eng.canvas)


def emu_pfn_bwd(eng, n_img, cache):
    bn = "pillar_vfe.pfn_layers.0.norm"
    dc = PI.act_to_nchw(eng.d_canvas, eng.d_canvas.n_cap)[:n_img]
    vc = cache["vc"]
    a = vc[:, 0].long()
    idx = (vc[:, 1] + vc[:, 2] * eng.nx + vc[:, 3]).long()
    dpf = dc.reshape(n_img, 64, eng.ny * eng.nx)[a, :, idx]
    dr = torch.zeros_like(cache["r"])
    dr.scatter_(1, cache["arg"].unsqueeze(1), dpf.unsqueeze(1))
    dy = dr * (cache["r"] > 0).float()
    dlin, dg, db = BO.bn_bwd(dy, cache["bc"])
    eng.G[bn + ".weight"].copy_(dg)
    eng.G[bn + ".bias"].copy_(db)
    eng.G["pillar_vfe.pfn_layers.0.linear.weight"].copy_(torch.einsum("msc,msf->cf", dlin, cache["feats"]))


def emu_heads_pack(eng, o):
    n = o["n_sc"]
    g = torch.cat([t[:n] for t in eng.head_grad], 1)                       # (n, tot, H, W)
    tot = g.shape[1]
    full = torch.zeros(eng.g_pf.n_cap, 64, o["H"], o["W"])
    full[:n, :tot] = g
    eng.g_pf.t.zero_()
    PI.nchw_to_act(full, eng.g_pf)
    c0 = 0
    for h, cn in zip(eng.head_mods, eng.head_cn):
        eng.G[h + ".bias"].copy_(g[:, c0:c0 + cn].sum((0, 2, 3)))
        c0 += cn


def run_train_plan(eng, args, batch, loss_grad_fn):
    """Forward + backward through the launch plan on the CPU.  Returns (head outputs, eng.G)."""
    record_len = [int(v) for v in batch["record_len"]]
    n_img, n_sc = sum(record_len), len(record_len)
    if "dx_in" not in eng.lvl[0] and len(eng.levels) > 1:
        eng._alloc_dx_in()
    plan = eng.build_train_ops(tuple(record_len))
    affine = O.normalize_pairwise_tfm(batch["pairwise_t_matrix"], eng.ny, eng.nx, float(args["voxel_size"][0]))
    method = args.get("fusion_method", "att")
    cache = {}

    def run(ops):
        for kind, o in ops:
            if kind == "conv":
                run_conv(eng, o["desc"])
            elif kind == "wgrad":
                emu_wgrad(eng, o)
            elif kind == "bn_fwd":
                emu_bn_fwd(eng, o)
            elif kind == "bn_apply":
                emu_bn_apply(eng, o)
            elif kind == "bn_bwd":
                emu_bn_bwd(eng, o)
            elif kind == "permute":
                emu_permute(eng, o)
            elif kind == "permute_batch":
                for q in o["jobs"]:
                    emu_permute(eng, q)
            elif kind == "pack_all":
                emu_pack_all(eng)
            elif kind == "head_bias":
                c0 = 0
                for h, cn in zip(eng.head_mods, eng.head_cn):
                    eng.head_bias[c0:c0 + cn].copy_(eng.P[h + ".bias"])
                    c0 += cn
            elif kind == "pfn_fwd":
                emu_pfn_fwd(eng, args, batch, n_img, cache)
            elif kind == "pfn_bwd":
                emu_pfn_bwd(eng, n_img, cache)
            elif kind == "affine":
                pass
            elif kind == "fuse":
                emu_fuse(eng, o, record_len, affine, method)
            elif kind == "fuse_bwd":
                emu_fuse_bwd(eng, o, record_len, affine, method)
            elif kind == "zero_grads":
                eng.wgflat.zero_(); eng.gflat.zero_(); eng.red_b.zero_(); eng.head_dbias.zero_()
            elif kind == "zero_fwd":
                eng.red_f.zero_()
            elif kind == "heads_pack":
                emu_heads_pack(eng, o)
            elif kind == "bucket":
                pass
            else:
                raise RuntimeError("unknown op " + kind)

    run(plan["fwd"])
    out = {name: t[:n_sc].clone() for name, t in zip(eng.head_names, eng.head_out)}
    grads = loss_grad_fn(out)
    for t, name in zip(eng.head_grad, eng.head_names):
        t[:n_sc].copy_(grads[name])
    run(plan["bwd"])
    return out, eng.G
