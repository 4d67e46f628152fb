"""What a plain streaming kernel reaches on this box at the training step's tensor sizes (calibration for the BatchNorm
kernels): torch copy / add / column-sum vs cb_bn_stats / cb_bn_apply / cb_bn_bwd_reduce / cb_bn_bwd_apply (development aid)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coalign_b200 import _lib  # noqa: E402

L = _lib.load()
BF16 = torch.bfloat16
sp = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=2 if os.environ.get("EXP_BW_SIZES") else 6):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


SIZES = ((20, 100, 352, 64), (20, 50, 176, 128), (20, 25, 88, 256))
if os.environ.get("EXP_BW_SIZES"):
    SIZES = tuple(SIZES[int(i)] for i in os.environ["EXP_BW_SIZES"].split(","))
for n, H, W, c in SIZES:
    rows = n * (H + 2) * (W + 2)
    mb = rows * c * 2 / 1e6
    a, b, y = (torch.randn(rows, c, device="cuda").to(BF16) for _ in range(3))
    o, o2 = torch.zeros_like(a), torch.zeros_like(a)
    m = _lib.Map()
    m.n_img, m.Hp, m.Wp, m.c_total, m.c_mod = n, H + 2, W + 2, c, c
    m.y_mode, m.y_pitch = 0, c
    sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    scale, shift, mean, inv, gamma, dg, db = (torch.ones(c, device="cuda") for _ in range(7))
    res = {}
    res["copy"] = (2, timeit(lambda: o.copy_(a)))
    res["add"] = (3, timeit(lambda: torch.add(a, b, out=o)))
    res["colsum(f32 acc)"] = (1, timeit(lambda: a.sum(0, dtype=torch.float32)))
    res["sum(all) torch"] = (1, timeit(lambda: a.sum(dtype=torch.float32)))      # read-only reference (same dirty-L2 start)
    res["bn_stats"] = (1, timeit(lambda: L.cb_bn_stats(a.data_ptr(), 0, C.byref(m), sums.data_ptr(), sp)))
    res["bn_apply"] = (2, timeit(lambda: L.cb_bn_apply(a.data_ptr(), 0, scale.data_ptr(), shift.data_ptr(), None, 0, None,
                                                       None, None, c, 0, 1, C.byref(m), o.data_ptr(), 0, sp)))
    res["bn_apply+res"] = (3, timeit(lambda: L.cb_bn_apply(a.data_ptr(), 0, scale.data_ptr(), shift.data_ptr(), None, 0, None,
                                                           None, b.data_ptr(), c, 0, 1, C.byref(m), o.data_ptr(), 0, sp)))
    res["bwd_reduce(zmask)"] = (2, timeit(lambda: L.cb_bn_bwd_reduce(a.data_ptr(), 0, None, 0, 1, b.data_ptr(), 0,
                                                                      mean.data_ptr(), inv.data_ptr(), scale.data_ptr(),
                                                                      shift.data_ptr(), C.byref(m), sums.data_ptr(), sp)))
    res["bwd_reduce(y)"] = (3, timeit(lambda: L.cb_bn_bwd_reduce(a.data_ptr(), 0, y.data_ptr(), 0, 1, b.data_ptr(), 0,
                                                                  mean.data_ptr(), inv.data_ptr(), None, None, C.byref(m),
                                                                  sums.data_ptr(), sp)))
    res["bwd_apply(zmask)"] = (3, timeit(lambda: L.cb_bn_bwd_apply(a.data_ptr(), 0, None, 0, 1, b.data_ptr(), 0, mean.data_ptr(),
                                                                    inv.data_ptr(), gamma.data_ptr(), scale.data_ptr(),
                                                                    shift.data_ptr(), sums.data_ptr(), float(n * H * W),
                                                                    C.byref(m), o.data_ptr(), 0, None, 0, dg.data_ptr(),
                                                                    db.data_ptr(), sp)))
    res["bwd_apply(y,dsum)"] = (5, timeit(lambda: L.cb_bn_bwd_apply(a.data_ptr(), 0, y.data_ptr(), 0, 1, b.data_ptr(), 0,
                                                                     mean.data_ptr(), inv.data_ptr(), gamma.data_ptr(), None,
                                                                     None, sums.data_ptr(), float(n * H * W), C.byref(m),
                                                                     o.data_ptr(), 0, o2.data_ptr(), 0, dg.data_ptr(),
                                                                     db.data_ptr(), sp)))
    print(f"[{n}x{H}x{W}x{c}] {mb:.1f} MB per tensor")
    for k, (nt, us) in res.items():
        print(f"  {k:20s} {us:7.1f} us  {nt * mb / us * 1e3:7.0f} GB/s ({nt} tensors)", flush=True)
