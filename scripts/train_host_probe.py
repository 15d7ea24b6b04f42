"""Is the cfg-5 training step bound by the host's launch rate or by the GPU?  (bf16 autocast: no scaler sync)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turtlevsr_b200.archs import create_video_model
from turtlevsr_b200.configs import shipped
from turtlevsr_b200.training import TrainStep
opt = shipped("Turtle_Derain"); torch.manual_seed(10)
net = create_video_model(opt).cuda()
ts = TrainStep(net, dict(lr=4e-4, weight_decay=0, betas=[0.9, 0.99]), amp="bf16")
g = torch.Generator().manual_seed(1)
lq = torch.rand(2, 5, 3, 256, 256, generator=g).cuda(); gt = torch.rand(2, 5, 3, 256, 256, generator=g).cuda()
for _ in range(3):
    ts.step(lq, gt)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); ts.step(lq, gt); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host issue {1e3*(t1-t0):.1f} ms, GPU drained after another {1e3*(t2-t1):.1f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    ts.step(lq, gt); torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = {}
for e in evs:
    d = agg.setdefault(e.name, [0.0, 0])
    d[0] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
    d[1] += 1
tot = sum(v[0] for v in agg.values()) / 1e3
print(f"sum of device kernel time {tot:.1f} ms over {sum(v[1] for v in agg.values())} kernels")
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"  {name[:86]:86s} {t/1e3:8.2f} ms  x{n}")
print("full names of the ATen elementwise / reduce fallbacks:")
for name, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
    if "elementwise_kernel" in name or "reduce_kernel" in name:
        print(f"  {t/1e3:7.2f} ms x{n}: {name[:400]}")
