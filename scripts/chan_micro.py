"""Isolated timing of the channel-attention tail (softmax over the Gram partials, fold into W_out) at the shapes of a
720p frame (CUDA events; every launch reads its own partial set so L2 does not keep them).
  python scripts/chan_micro.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from turtlevsr_b200.capi import call

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
st = torch.cuda.current_stream().cuda_stream
ch = 64
for heads, S in [(1, 1), (2, 1), (4, 1), (8, 1), (4, 3), (8, 4), (2, 3), (1, 3)]:
    c = heads * ch
    nsplit = 296 // heads
    nset = int(os.environ.get("NSET", "6"))
    sets = [(torch.randn(S, nsplit, heads, ch, ch, device="cuda"), torch.rand(S, nsplit, c, device="cuda") + 1,
             torch.rand(S, nsplit, c, device="cuda") + 1) for _ in range(nset)]
    flags = torch.zeros(S, dtype=torch.int32, device="cuda")
    temp = torch.ones(heads, device="cuda")
    Pm = torch.empty(heads, ch, S * ch, device="cuda")
    inv = torch.empty(S, c, device="cuda")
    Wo = torch.randn(c, c, device="cuda")
    M = torch.empty(c, S * c, device="cuda", dtype=torch.float16)

    def softmax(i):
        g, a, b = sets[i % nset]
        call("turtle_chan_softmax", g.data_ptr(), a.data_ptr(), b.data_ptr(), flags.data_ptr(), temp.data_ptr(), S, nsplit,
             heads, ch, Pm.data_ptr(), inv.data_ptr(), st)

    def fold(i):
        call("turtle_chan_fold", Pm.data_ptr(), Wo.data_ptr(), S, heads, ch, M.data_ptr(), 2, st)

    out = []
    for fn in (softmax, fold):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        # `reps` launches captured into one CUDA graph: the replay shows the device time, not the host's launch rate
        gs = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(gs):
            st = gs.cuda_stream
            with torch.cuda.graph(g, stream=gs):
                for i in range(reps):
                    fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        st = torch.cuda.current_stream().cuda_stream
        out.append(e0.elapsed_time(e1) * 1e3 / (5 * reps))
    print(f"heads={heads} S={S} nsplit={nsplit}: softmax {out[0]:6.2f} us  fold {out[1]:6.2f} us", flush=True)
