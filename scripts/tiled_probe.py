"""Tiled 720p inference (tile 320, overlap 128 = 24 tiles, INF:172-246 defaults of the reference's app): ms/frame
tile by tile (eager / CUDA graphs) vs all tiles as one batch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model
import turtlevsr_b200.tiling as tl
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
clip = torch.rand(6, 1, 3, 720, 1280, generator=g).to(dev)
for name, graphs, batch in [("tile by tile, eager", False, False), ("tile by tile, graphs", True, False),
                            ("batched tiles, eager", False, True), ("batched tiles, graphs", True, True)]:
    net, _ = build_model("tf32", dev)
    net.enable_cuda_graphs(graphs, max_graphs=160)            # 24 tile histories x 6 ring states
    dk = dv = None
    n_warm, n = 14, 6
    with torch.no_grad():
        for j in range(n_warm + n):
            if j == n_warm:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            out, dk, dv = tl.run_inference_patched(clip[(j - 1) % 6 if j else 0], clip[j % 6], net, dev, 320, 128,
                                                   prev_patch_dict_k=dk, prev_patch_dict_v=dv, model_type="t1",
                                                   batch_tiles=batch)
        torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / n
    print(f"{name:24s}: {ms:7.1f} ms/frame ({len(dk) - (1 if batch else 0)} tiles), mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del net, dk, dv
    torch.cuda.empty_cache()
