import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from turtlevsr_b200.capi import call
from gpu_util import stream
torch.manual_seed(0)
for (P,heads,nsplit) in [(640,1,1),(700,2,5),(4096,4,7)]:
    c=heads*64
    x=torch.randn(P,3*c,device='cuda').half()
    g=torch.full((nsplit,heads,64,64),float('nan'),device='cuda'); sqq=torch.zeros(nsplit,c,device='cuda'); sqk=torch.zeros(nsplit,c,device='cuda')
    call("turtle_chan_gram", x.data_ptr(), 3*c, 64, x.data_ptr()+2*c, 3*c, 64, P, heads, 64, nsplit, g.data_ptr(), sqq.data_ptr(), sqk.data_ptr(), 2, stream())
    torch.cuda.synchronize()
    G=g.sum(0); xd=x.double()
    q=xd[:, :c].reshape(P,heads,64); k=xd[:, c:2*c].reshape(P,heads,64)
    want=torch.einsum('phi,phj->hij', q,k)
    err=(G.double()-want).abs()
    print(P,heads,nsplit,'G maxerr',err.max().item(),'scale',want.abs().max().item(), 'sqq err',(sqq.sum(0).double()-(xd[:,:c]**2).sum(0)).abs().max().item(), 'sqk err',(sqk.sum(0).double()-(xd[:,c:2*c]**2).sum(0)).abs().max().item())
