import sys, torch, ctypes as C
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from gpu_util import gemm
from turtlevsr_b200 import capi
P,Cin,Cout=1000,64,128
A=torch.randn(P,Cin,device='cuda'); W=torch.randn(Cout,Cin,device='cuda')/8
out=torch.zeros(P,Cout,device='cuda')
for kw in [dict(), dict(res=torch.randn(P,Cout,device='cuda')), dict(bias=torch.randn(Cout,device='cuda'))]:
    try:
        out.zero_()
        gemm([(A,0,Cin)],Cin,W,P,Cout,mode=1,out=out,ldo=Cout,**kw)
        torch.cuda.synchronize()
        want=A@W.t()
        if 'res' in kw: want=want+kw['res']
        if 'bias' in kw: want=want+kw['bias']
        print(list(kw), 'ok err', (out-want).abs().max().item())
    except Exception as e:
        print(list(kw), 'EXC', repr(e)[:300]); break
