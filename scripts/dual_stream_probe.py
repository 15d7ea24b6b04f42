"""Probe: frames/s with 1 vs 2 (vs 3) independent clips in flight on one GPU (one stream + engine + rings per clip)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model
dev = torch.device("cuda")
NC = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 16
nets = [build_model("tf32", dev)[0] for _ in range(NC)]
streams = [torch.cuda.Stream() for _ in range(NC)]
g = torch.Generator().manual_seed(0)
clips = [torch.rand(4, 1, 3, 720, 1280, generator=g).to(dev) for _ in range(NC)]
ks = [None] * NC; vs = [None] * NC; outs = [None] * NC
def frame(c, j):
    cl = clips[c]
    x = torch.stack([cl[(j - 1) % 4 if j else 0], cl[j % 4]], 1)
    outs[c], ks[c], vs[c] = nets[c](x, ks[c], vs[c])
with torch.no_grad():
    for mode in ("serial", "interleaved"):
        ks = [None] * NC; vs = [None] * NC
        torch.cuda.synchronize()
        for j in range(4 + K):
            if j == 4:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            for c in range(NC):
                if mode == "serial":
                    frame(c, j)
                else:
                    with torch.cuda.stream(streams[c]):
                        frame(c, j)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{mode}: {NC} clips x {K} frames in {dt*1e3:.1f} ms -> {NC*K/dt:.2f} frames/s", flush=True)
        if mode == "serial":
            ref = [o.clone() for o in outs]
        else:
            print("max |interleaved - serial| =", max((a - b).abs().max().item() for a, b in zip(outs, ref)))
