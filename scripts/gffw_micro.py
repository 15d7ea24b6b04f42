"""Fused GatedFeedForward kernel vs the unfused schedule it replaces, each hot 720p shape alone (CUDA events, rotating
buffers larger than L2).  python scripts/gffw_micro.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from turtlevsr_b200 import capi
from turtlevsr_b200.capi import GemmArgs, call

dev = torch.device("cuda")
s = lambda: torch.cuda.current_stream().cuda_stream
PEAK = 6557.1


def gemm16(A16, K, W16, out, ldo, P, Cout, o16, res=None):
    a = GemmArgs()
    a.mode, a.P, a.Cout, a.nseg, a.segw = capi.TF32, P, Cout, 1, K
    a.A[0], a.lda[0] = A16.data_ptr(), K
    a.Wt = W16.data_ptr()
    a.out, a.ldo = out.data_ptr(), ldo
    a.a_dtype, a.out_dtype = 1, 1 if o16 else 0
    if res is not None:
        a.res, a.ldres = res.data_ptr(), Cout
    call("turtle_gemm", C.byref(a), s())


def timed(fn, nrot, reps=3):
    for i in range(nrot):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(nrot):
            fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * nrot) * 1e3      # us


for c, H, W in [(64, 736, 1280), (128, 368, 640), (256, 184, 320)]:
    hid = int(2.5 * c)
    P = H * W
    nrot = max(2, int(400e6 / (P * c * 12)) + 1)
    g = torch.Generator().manual_seed(c)
    w_in = (torch.randn(2 * hid, c, generator=g) / c ** 0.5).half().to(dev)
    wdw = (torch.randn(2 * hid, 9, generator=g) / 3)
    w_out = (torch.randn(c, hid, generator=g) / hid ** 0.5).half().to(dev)
    taps_f = wdw.reshape(2, hid // 32, 32, 9).permute(1, 0, 3, 2).contiguous().half().to(dev)
    taps_u = wdw.t().contiguous().half().to(dev)
    ln_w, ln_b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    xn = [torch.randn(P, c, device=dev).half() for _ in range(nrot)]
    x = [torch.randn(P, c, device=dev) for _ in range(nrot)]
    ln = [torch.empty(P, c, device=dev, dtype=torch.float16) for _ in range(nrot)]
    t16 = torch.empty(P, 2 * hid, device=dev, dtype=torch.float16)
    g16 = torch.empty(P, hid, device=dev, dtype=torch.float16)

    def fused(i):
        call("turtle_gffw_fused", xn[i].data_ptr(), w_in.data_ptr(), taps_f.data_ptr(), w_out.data_ptr(), x[i].data_ptr(),
             ln[i].data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), 1, H, W, c, hid, s())

    def unfused(i):
        gemm16(xn[i], c, w_in, t16, 2 * hid, P, 2 * hid, True)
        call("turtle_dwconv3x3", t16.data_ptr(), 2 * hid, taps_u.data_ptr(), None, g16.data_ptr(), hid, 1, H, W, 2 * hid, 2, 0,
             1, 2, s())
        gemm16(g16, hid, w_out, x[i], c, P, c, False, res=x[i])

    def tail(i):                # project_in GEMM + (depthwise + gate + project_out) kernel
        gemm16(xn[i], c, w_in, t16, 2 * hid, P, 2 * hid, True)
        call("turtle_gffw_tail", t16.data_ptr(), taps_f.data_ptr(), w_out.data_ptr(), x[i].data_ptr(), ln[i].data_ptr(),
             ln_w.data_ptr(), ln_b.data_ptr(), 1, H, W, c, hid, s())

    def tail_only(i):
        call("turtle_gffw_tail", t16.data_ptr(), taps_f.data_ptr(), w_out.data_ptr(), x[i].data_ptr(), ln[i].data_ptr(),
             ln_w.data_ptr(), ln_b.data_ptr(), 1, H, W, c, hid, s())

    if os.environ.get("GFFW_ONCE"):          # one launch per shape (for ncu)
        (tail_only if os.environ["GFFW_ONCE"] == "tail" else fused)(0)
        torch.cuda.synchronize()
        continue
    tt, tto = timed(tail, nrot), timed(tail_only, nrot)
    print(f"GFFW c={c} {H}x{W}: project_in GEMM + gffw_tail {tt:8.1f} us (tail kernel alone {tto:8.1f} us)", flush=True)
    tf, tu = timed(fused, nrot), timed(unfused, nrot)
    by = P * c * (2 + 4 + 4 + 2)
    fl = 2 * P * c * 3 * hid
    print(f"GFFW c={c} {H}x{W}: fused {tf:8.1f} us ({by / tf / 1e3:6.0f} GB/s = {by / tf / 1e3 / PEAK:.2f} of copy peak, "
          f"{fl / tf / 1e6:6.0f} TF/s nominal)   unfused chain {tu:8.1f} us   speed-up {tu / tf:.2f}x", flush=True)
