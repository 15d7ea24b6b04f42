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
ka = prof.key_averages()
tot = sum(e.device_time_total for e in ka) / 1e3
print(f"sum of device kernel time {tot:.1f} ms")
for e in sorted(ka, key=lambda e: -e.device_time_total)[:14]:
    print(f"  {e.key[:70]:70s} {e.device_time_total/1e3:8.2f} ms  x{e.count}")
