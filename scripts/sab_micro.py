"""StateAlignBlock aggregation at the three decoder shapes of a 720p frame: CUDA-core quad kernel vs the tensor-core
path (dense key-box contraction + far top-k gather).  CUDA events, fp16 output, random selections (top-k keys far away).
  python scripts/sab_micro.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turtlevsr_b200 import capi
from turtlevsr_b200.capi import call

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
st = torch.cuda.current_stream().cuda_stream
Hg, Wg, F_, D = 46, 80, 3, 128
N = Hg * Wg
g = torch.Generator(device="cuda").manual_seed(0)
q = torch.nn.functional.normalize(torch.randn(N, D, device="cuda", generator=g), dim=-1)
k = torch.nn.functional.normalize(torch.randn(F_, N, D, device="cuda", generator=g), dim=-1)
idx = torch.empty(F_, N, capi.SAB_SLOTS, dtype=torch.int32, device="cuda")
wgt = torch.empty(F_, N, capi.SAB_SLOTS, device="cuda")
tau = torch.tensor([0.83], device="cuda")
call("turtle_sab_select", q.data_ptr(), k.data_ptr(), N * D, F_, Hg, Wg, D, tau.data_ptr(), 0, idx.data_ptr(), wgt.data_ptr(), 0, st)
wsp = torch.empty(capi.load().turtle_sab_aggregate_tc_workspace(F_, Hg, Wg) // 4, device="cuda")
tot = [0.0, 0.0, 0.0]
for ws, c in [(4, 256), (8, 128), (16, 64)]:
    Dv = ws * ws * c
    V = torch.randn(F_, N, Dv, device="cuda", generator=g)
    V = ((V.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    y0 = torch.empty(F_, Hg * ws, Wg * ws, c, device="cuda", dtype=torch.float16)
    y1 = torch.empty_like(y0)
    y2 = torch.empty_like(y0)
    V16 = V.half()

    def old():
        call("turtle_sab_aggregate", idx.data_ptr(), wgt.data_ptr(), V.data_ptr(), N * Dv, y0.data_ptr(), F_, Hg, Wg, ws, c, 0, 2, st)

    def new():
        call("turtle_sab_aggregate_tc", idx.data_ptr(), wgt.data_ptr(), V.data_ptr(), 0, N * Dv, y1.data_ptr(), F_, Hg, Wg, ws, c, 2,
             wsp.data_ptr(), st)

    def new16():
        call("turtle_sab_aggregate_tc", idx.data_ptr(), wgt.data_ptr(), V16.data_ptr(), 1, N * Dv, y2.data_ptr(), F_, Hg, Wg, ws, c, 2,
             wsp.data_ptr(), st)

    res = []
    for fn in (old, new, new16):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 1e3 / reps)
    tot[0] += res[0]; tot[1] += res[1]; tot[2] += res[2]
    by = F_ * N * Dv * 4 + y0.numel() * 2
    err = (y0.float() - y1.float()).abs().max().item()
    err2 = (y0.float() - y2.float()).abs().max().item()
    print(f"ws={ws:2d} c={c:3d} Dv={Dv:5d}: quad {res[0]:7.1f} us  tensor-core path: TF32 rows {res[1]:7.1f} us, fp16 rows {res[2]:7.1f} us  "
          f"(HBM floor {by / 6557e3:6.1f} us)  max|diff| {err:.2e} / {err2:.2e}", flush=True)
print(f"per frame: quad {tot[0]:.0f} us, tensor-core path {tot[1]:.0f} us (TF32 rows), {tot[2]:.0f} us (fp16 rows)")
