import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turtlevsr_b200.capi import call
st = torch.cuda.current_stream().cuda_stream
B, C, H, W = 1, 64, 9, 14
Co = C // 2
# x[b,y,x,c] = small distinct values; identity taps; out = gelu(u1)*u2
x = torch.zeros(B, H, W, C)
for c in range(C):
    x[..., c] = (c + 1) / 8.0
x[0, 3, 5, :] += 1.0
x = x.half().cuda()
w9 = torch.zeros(9, C); w9[4] = 1.0
w9 = w9.half().cuda()
out = torch.full((B, H, W, Co), float("nan"), device="cuda", dtype=torch.float16)
call("turtle_dwconv3x3", x.data_ptr(), C, w9.data_ptr(), None, out.data_ptr(), Co, B, H, W, C, 2, 0, 1, 2, st)
torch.cuda.synchronize()
want = torch.nn.functional.gelu(x[..., :Co].float()) * x[..., Co:].float()
print("max err", (out.float() - want).abs().max().item())
print("got  [0,2,2,:8]", out[0, 2, 2, :8].float().tolist())
print("want [0,2,2,:8]", want[0, 2, 2, :8].tolist())
print("got  [0,3,5,:8]", out[0, 3, 5, :8].float().tolist())
print("want [0,3,5,:8]", want[0, 3, 5, :8].tolist())
print("got  [0,0,0,:]", out[0, 0, 0, :].float().tolist())
print("want [0,0,0,:]", want[0, 0, 0, :].tolist())
# plain variant for comparison
out0 = torch.full((B, H, W, C), float("nan"), device="cuda", dtype=torch.float16)
call("turtle_dwconv3x3", x.data_ptr(), C, w9.data_ptr(), None, out0.data_ptr(), C, B, H, W, C, 0, 0, 1, 2, st)
print("plain max err", (out0.float() - x.float()).abs().max().item())
