import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from turtlevsr_b200.capi import call
from gpu_util import stream
torch.manual_seed(0)
P,heads,nsplit=640,1,1
c=heads*64
x=torch.randn(P,3*c,device='cuda')
g=torch.full((nsplit,heads,64,64),float('nan'),device='cuda'); sqq=torch.full((nsplit,c),float('nan'),device='cuda'); sqk=torch.full((nsplit,c),float('nan'),device='cuda')
call("turtle_chan_gram", x.data_ptr(), 3*c, 64, x.data_ptr()+4*c, 3*c, 64, P, heads, 64, nsplit, g.data_ptr(), sqq.data_ptr(), sqk.data_ptr(), 1, stream())
torch.cuda.synchronize()
print('g nan frac', g.isnan().float().mean().item(), 'zeros frac', (g==0).float().mean().item(), 'sqq', sqq[0,:4].tolist(), 'sqk', sqk[0,:4].tolist())
print(g[0,0,:2,:8])
