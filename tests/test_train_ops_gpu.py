"""GPU op-level tests of the training kernels through the C ABI: each kernel against a plain PyTorch
fp32/fp64 evaluation of the same op on the SAME bf16-rounded operands, so the tolerances are those of fp32 accumulation
order, not of bf16 rounding."""
import ctypes as C

import numpy as np
import pytest
import torch

from coalign_b200 import _lib

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def lib():
    return _lib.load(check_device=True)


def sp():
    return torch.cuda.current_stream().cuda_stream


def pf_from_nchw(x, pitch=None):
    """(n,C,H,W) float -> bf16 PF rows [n*(H+2)*(W+2)][pitch] with zero halo."""
    n, c, h, w = x.shape
    pitch = pitch or c
    t = torch.zeros(n, h + 2, w + 2, pitch, dtype=BF16, device=x.device)
    t[:, 1:h + 1, 1:w + 1, :c] = x.permute(0, 2, 3, 1).to(BF16)
    return t.reshape(-1, pitch).contiguous()


def nchw_from_pf(t, n, c, h, w):
    return t.float().reshape(n, h + 2, w + 2, -1)[:, 1:h + 1, 1:w + 1, :c].permute(0, 3, 1, 2).contiguous()


def conv_wgrad_desc(dz, x, dw, n, H, W, cin, cout, simt=False, k_splits=0):
    """3x3/s1: dw fp32 [cout][9*cin] (tap-major K), units = 128 dz channels x up to four 64-channel (tap, block) boxes."""
    d = _lib.WgradDesc()
    Wp = W + 2
    d.dz_ptr, d.dz_lo_ptr, d.dz_pitch, d.k_splits = dz.data_ptr(), None, cout, k_splits
    d.rows_total = n * (H + 2) * Wp
    d.x_ptr[0], d.x_rows[0], d.x_pitch[0] = x.data_ptr(), x.shape[0], cin
    d.dw = dw.data_ptr()
    boxes = [(t, cb) for t in range(9) for cb in range(cin // 64)]
    nu = 0
    for m0 in range(0, cout, 128):
        for i in range(0, len(boxes), 4):
            u = d.units[nu]
            u.m0, u.a_row_off, u.m_valid = m0, 0, min(128, cout - m0)
            grp = boxes[i:i + 4]
            u.n_boxes = len(grp)
            for j, (t, cb) in enumerate(grp):
                r, s = divmod(t, 3)
                u.box[j].row_off = (r - 1) * Wp + (s - 1)
                u.box[j].col, u.box[j].x_sel = cb * 64, 0
                u.box[j].out_ld = 9 * cin
                u.box[j].out_off = m0 * 9 * cin + t * cin + cb * 64
            nu += 1
    d.n_units = nu
    return d


@pytest.mark.parametrize("cin,cout,n,H,W", [(64, 64, 2, 12, 28), (128, 256, 3, 7, 11), (256, 128, 1, 25, 88), (384, 256, 1, 9, 20)])
def test_wgrad_tensor_core_matches_torch(cin, cout, n, H, W):
    """cb_wgrad (MN-major tcgen05, split-K, vector reductions) == torch.nn.grad.conv2d_weight on the same bf16 operands;
    also == the SIMT evaluation of the same descriptor."""
    from torch.nn.grad import conv2d_weight
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    x = torch.randn(n, cin, H, W, device="cuda", generator=g)
    dzv = torch.randn(n, cout, H, W, device="cuda", generator=g)
    xp, dzp = pf_from_nchw(x), pf_from_nchw(dzv)
    ref = conv2d_weight(xp.float().reshape(n, H + 2, W + 2, cin)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).double(),
                        (cout, cin, 3, 3),
                        dzp.float().reshape(n, H + 2, W + 2, cout)[:, 1:-1, 1:-1].permute(0, 3, 1, 2).double(), padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(cout, 9 * cin).float()           # [cout][tap][cin]
    for ks in (0, 1, 3):
        dw = torch.zeros(cout, 9 * cin, device="cuda")
        d = conv_wgrad_desc(dzp, xp, dw, n, H, W, cin, cout, k_splits=ks)
        descs = [d]
        if d.n_units > _lib.CB_WGRAD_MAX_UNITS:
            pytest.skip("too many units for one launch")
        _lib.check(lib().cb_wgrad(C.byref(d), 0, sp()), "cb_wgrad")
        torch.cuda.synchronize()
        err = (dw - ref).abs().max().item()
        scale = ref.abs().max().item()
        assert err <= 2e-4 * scale + 1e-3, (ks, err, scale)
    dw2 = torch.zeros(cout, 9 * cin, device="cuda")
    d2 = conv_wgrad_desc(dzp, xp, dw2, n, H, W, cin, cout)
    _lib.check(lib().cb_wgrad_simt(C.byref(d2), sp()), "cb_wgrad_simt")
    torch.cuda.synchronize()
    assert (dw2 - ref).abs().max().item() <= 2e-4 * ref.abs().max().item() + 1e-3


def make_map(n, H, W, c_total, c_mod=None, y_mode=0, y_pitch=None, y_ch_off=0, up_k=0, y_Hp=0, y_Wp=0, plane_rows=0):
    m = _lib.Map()
    m.n_img, m.Hp, m.Wp, m.c_total, m.c_mod = n, H + 2, W + 2, c_total, c_mod or c_total
    m.y_mode, m.y_pitch, m.y_ch_off, m.up_k, m.y_Hp, m.y_Wp, m.y_plane_rows = (y_mode, y_pitch or c_total, y_ch_off, up_k, y_Hp,
                                                                              y_Wp, plane_rows)
    return m


def test_train_mode_batchnorm_forward_backward_pf():
    """cb_bn_stats / finalize / apply / bwd_reduce / bwd_apply on a PF tensor == F.batch_norm(training=True) + ReLU and its
    autograd, including the running-stat update."""
    torch.manual_seed(0)
    n, c, H, W = 3, 64, 9, 14
    z = (torch.randn(n, c, H, W, device="cuda") * 2 + 0.5).to(BF16).float()
    res = torch.randn(n, c, H, W, device="cuda").to(BF16).float()
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda") * 0.3
    rm, rv = torch.randn(c, device="cuda"), torch.rand(c, device="cuda") + 0.5
    rm0, rv0 = rm.clone(), rv.clone()
    zp, rp = pf_from_nchw(z), pf_from_nchw(res)
    m = make_map(n, H, W, c)
    sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    scale, shift, mean, inv = (torch.empty(c, device="cuda") for _ in range(4))
    L = lib()
    _lib.check(L.cb_bn_stats(zp.data_ptr(), 0, C.byref(m), sums.data_ptr(), sp()))
    _lib.check(L.cb_bn_finalize(sums.data_ptr(), c, float(n * H * W), 1e-5, 0.1, gamma.data_ptr(), beta.data_ptr(),
                                rm.data_ptr(), rv.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                inv.data_ptr(), sp()))
    # the one-launch form (last CTA closes the statistics) must give the same numbers and leave its ticket counter at zero
    sums2 = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    rm2, rv2 = rm0.clone(), rv0.clone()
    sc2, sh2, mean2, inv2 = (torch.empty(c, device="cuda") for _ in range(4))
    for _ in range(2):                                                    # twice: the counter must reset itself
        sums2.zero_()
        rm2.copy_(rm0)
        rv2.copy_(rv0)
        _lib.check(L.cb_bn_stats_finalize(zp.data_ptr(), 0, C.byref(m), sums2.data_ptr(), ticket.data_ptr(), float(n * H * W),
                                          1e-5, 0.1, gamma.data_ptr(), beta.data_ptr(), rm2.data_ptr(), rv2.data_ptr(),
                                          sc2.data_ptr(), sh2.data_ptr(), mean2.data_ptr(), inv2.data_ptr(), sp()))
    torch.cuda.synchronize()
    assert int(ticket.item()) == 0
    for a, b in ((sc2, scale), (sh2, shift), (mean2, mean), (inv2, inv), (rm2, rm), (rv2, rv)):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
    y = torch.zeros_like(zp)
    _lib.check(L.cb_bn_apply(zp.data_ptr(), 0, scale.data_ptr(), shift.data_ptr(), None, 0, None, None, rp.data_ptr(), c, 0, 1,
                             C.byref(m), y.data_ptr(), 0, sp()))
    rm_t, rv_t = rm0.clone(), rv0.clone()
    yt = torch.relu(torch.nn.functional.batch_norm(z, rm_t, rv_t, gamma, beta, True, 0.1, 1e-5) + res)
    torch.cuda.synchronize()
    got = nchw_from_pf(y, n, c, H, W)
    assert (got - yt).abs().max().item() < 2e-2 * yt.abs().max().item()                     # bf16 store of y
    assert torch.allclose(rm, rm_t, atol=1e-5) and torch.allclose(rv, rv_t, rtol=1e-5, atol=1e-5)
    assert y.reshape(n, H + 2, W + 2, c)[:, 0].abs().max().item() == 0      # halo untouched
    # backward with the kernel's own y as the ReLU mask
    dy = torch.randn(n, c, H, W, device="cuda").to(BF16).float()
    dyp = pf_from_nchw(dy)
    bs = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    _lib.check(L.cb_bn_bwd_reduce(dyp.data_ptr(), 0, y.data_ptr(), 0, 1, zp.data_ptr(), 0, mean.data_ptr(), inv.data_ptr(),
                                  None, None, C.byref(m), bs.data_ptr(), sp()))
    dz, dsum = torch.zeros_like(zp), torch.zeros_like(zp)
    dg, db = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    _lib.check(L.cb_bn_bwd_apply(dyp.data_ptr(), 0, y.data_ptr(), 0, 1, zp.data_ptr(), 0, mean.data_ptr(), inv.data_ptr(),
                                 gamma.data_ptr(), None, None, bs.data_ptr(), float(n * H * W), C.byref(m), dz.data_ptr(), 0,
                                 dsum.data_ptr(), 0, dg.data_ptr(), db.data_ptr(), sp()))
    mask = (got > 0).float()
    torch.cuda.synchronize()
    # reference gradients with our mask: recompute explicitly instead of relying on the line above
    xh = (z - z.mean((0, 2, 3), keepdim=True)) / torch.sqrt(z.var((0, 2, 3), unbiased=False, keepdim=True) + 1e-5)
    dyr = dy * mask
    db_ref, dg_ref = dyr.sum((0, 2, 3)), (dyr * xh).sum((0, 2, 3))
    cnt = n * H * W
    inv_ref = 1.0 / torch.sqrt(z.var((0, 2, 3), unbiased=False) + 1e-5)
    dz_ref = (gamma * inv_ref).view(1, -1, 1, 1) * (dyr - db_ref.view(1, -1, 1, 1) / cnt - xh * dg_ref.view(1, -1, 1, 1) / cnt)
    assert torch.allclose(db, db_ref, rtol=1e-4, atol=1e-3) and torch.allclose(dg, dg_ref, rtol=1e-4, atol=1e-3)
    assert (nchw_from_pf(dz, n, c, H, W) - dz_ref).abs().max().item() < 2e-2 * dz_ref.abs().max().item() + 1e-3
    assert (nchw_from_pf(dsum, n, c, H, W) - dyr).abs().max().item() == 0
    assert dz.reshape(n, H + 2, W + 2, c)[:, :, 0].abs().max().item() == 0


def test_adam_step_matches_torch_adam():
    torch.manual_seed(1)
    n = 100003
    p0 = torch.randn(n, device="cuda")
    p = p0.clone()
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-3, eps=1e-10, weight_decay=1e-4)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    for it in range(5):
        g = torch.randn(n, device="cuda") * (0.1 + it)
        ref.grad = g.clone()
        opt.step()
        _lib.check(lib().cb_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 2e-3, 0.9, 0.999, 1e-10, 1e-4, 1.0,
                                      step.data_ptr(), 1, sp()))
    torch.cuda.synchronize()
    assert int(step.item()) == 5
    assert torch.allclose(p, ref.detach(), rtol=1e-5, atol=1e-6), (p - ref.detach()).abs().max().item()


def test_pack_weight_and_permute():
    torch.manual_seed(2)
    cout, cin = 24, 16
    w = torch.randn(cout, cin, 3, 3, device="cuda")
    fwd = torch.zeros(cout, 2 * 9 * cin, dtype=BF16, device="cuda")
    _lib.check(lib().cb_pack_weight(w.data_ptr(), 1, cout, 9, cin, 0, cin * 9, 1, 9, fwd.data_ptr(), 2 * 9 * cin, 0, 9 * cin, sp()))
    ref = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    hi = ref.to(BF16)
    torch.cuda.synchronize()
    assert torch.equal(fwd[:, :9 * cin], hi) and torch.equal(fwd[:, 9 * cin:], (ref - hi.float()).to(BF16))
    dg = torch.zeros(cin, 9 * cout, dtype=BF16, device="cuda")
    _lib.check(lib().cb_pack_weight(w.data_ptr(), 1, cin, 9, cout, 0, 9, 1, cin * 9, dg.data_ptr(), 9 * cout, 0, 0, sp()))
    torch.cuda.synchronize()
    assert torch.equal(dg, w.permute(1, 2, 3, 0).reshape(cin, 9 * cout).to(BF16))
    packed = torch.randn(cout, 9 * cin, device="cuda")                      # [co][tap][ci] -> OIHW
    back = torch.empty(cout, cin, 3, 3, device="cuda")
    _lib.check(lib().cb_permute_f32(packed.data_ptr(), cout, cin, 1, 9, 9 * cin, 1, 0, cin, 0.5, back.data_ptr(), sp()))
    torch.cuda.synchronize()
    assert torch.allclose(back, 0.5 * packed.view(cout, 3, 3, cin).permute(0, 3, 1, 2))
