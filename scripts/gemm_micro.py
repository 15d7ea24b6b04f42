"""Isolated timing of the hot GEMM shapes of a 720p frame (CUDA events, rotating buffer sets larger than L2).
  python scripts/gemm_micro.py [reps]          ->  us per launch, GB/s, TF/s per shape"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turtlevsr_b200 import capi
from turtlevsr_b200.capi import GemmArgs, call

SHAPES = [  # (Cin, Cout, P, o16, res, ln)
    (256, 1280, 58880, 1, 0, 0), (256, 256, 58880, 0, 1, 1), (640, 256, 58880, 0, 1, 1), (256, 768, 58880, 1, 0, 0),
    (256, 128, 235520, 0, 1, 1), (128, 256, 235520, 1, 0, 0), (128, 640, 235520, 1, 0, 0),
    (128, 64, 942080, 0, 1, 1), (64, 128, 942080, 1, 0, 0), (64, 320, 942080, 1, 0, 0), (160, 64, 942080, 0, 1, 0),
    (512, 2560, 14720, 1, 0, 0), (1280, 512, 14720, 0, 1, 0), (512, 1536, 14720, 1, 0, 0),
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
st = torch.cuda.current_stream().cuda_stream
for si, (Cin, Cout, P, o16, res, ln) in enumerate(SHAPES):
    if only >= 0 and si != only:
        continue
    nset = 4
    sets = []
    for _ in range(nset):
        A = (torch.randn(P, Cin, device="cuda") * 0.5).half()
        out = torch.zeros(P, Cout, device="cuda", dtype=torch.float16 if o16 else torch.float32)
        xn = torch.zeros(P, Cout, device="cuda", dtype=torch.float16) if ln else None
        sets.append((A, out, xn))
    Wt = (torch.randn(Cout, Cin, device="cuda") / Cin ** 0.5).half()
    lw, lb = torch.ones(Cout, device="cuda"), torch.zeros(Cout, device="cuda")
    def launch(i):
        A, out, xn = sets[i % nset]
        a = GemmArgs()
        a.mode, a.im2col, a.P, a.Cout, a.nseg, a.segw = capi.TF32, 0, P, Cout, 1, Cin
        a.A[0], a.lda[0] = A.data_ptr(), Cin
        a.Wt = Wt.data_ptr()
        if res:
            a.res, a.ldres = out.data_ptr(), Cout
        a.out, a.ldo, a.store = out.data_ptr(), Cout, capi.STORE_PLAIN
        a.a_dtype, a.out_dtype = 1, o16
        if ln:
            a.ln_out, a.ld_ln, a.ln_w, a.ln_b = xn.data_ptr(), Cout, lw.data_ptr(), lb.data_ptr()
        call("turtle_gemm", C.byref(a), st)
    for i in range(3):
        launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        launch(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    by = P * Cin * 2 + Cout * Cin * 2 + P * Cout * ((2 if o16 else 4) + (4 if res else 0) + (2 if ln else 0))
    fl = 2.0 * P * Cin * Cout
    print(f"{si:2d} {Cin:5d}->{Cout:5d} @{P:7d} o16={o16} res={res} ln={ln}: {us:8.2f} us  {by/us/1e3:7.0f} GB/s  {fl/us/1e6:7.1f} TF/s", flush=True)
