"""Per-clip cached frame loop and clip sharding.

``run_clip`` is the canonical loop of the reference (VRM:110-129, INF:276-308): frame j is fed as
``stack([frame[j or j-1], frame[j]])`` together with the caches returned for frame j-1.  History is
strictly sequential inside a clip, so multi-GPU inference shards *clips* (VRM:162-164:
``idx % world_size == rank``) and needs no collective.
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch


@torch.no_grad()
def run_clip(net, clip: torch.Tensor, keep_outputs: bool = True):
    """clip [B,T,C,H,W] on the model's device -> (outputs [B,T,C,H',W'] or None, k_cache, v_cache)."""
    outs = []
    k = v = None
    T = clip.shape[1]
    for j in range(T):
        pre = clip[:, j if j == 0 else j - 1]
        x = torch.stack([pre, clip[:, j]], dim=1)
        o, k, v = net(x.float(), k, v)
        if keep_outputs:
            outs.append(o)
    return (torch.stack(outs, dim=1) if keep_outputs else None), k, v


def shard_clips(n_clips: int, rank: int, world_size: int) -> List[int]:
    """Clip indices owned by ``rank`` (same rule as VRM:162-164)."""
    return [i for i in range(n_clips) if i % world_size == rank]


@torch.no_grad()
def run_clip_streamed(net, clip_host: torch.Tensor, out_host: torch.Tensor = None, device=None, k=None, v=None,
                      prev: torch.Tensor = None):
    """The same loop for a clip that lives in (pinned) HOST memory: frame j+1 travels host->device on a copy stream while
    frame j is restored, and restored frame j-1 travels back on a third stream, so the PCIe traffic of
    ``inference.py``'s per-frame ``.to(device)`` / ``.cpu()`` (INF:276-308) disappears behind the kernels.

    clip_host [B,T,C,H,W] (pinned for true overlap); out_host [B,T,C,H',W'] pinned or None (allocated).
    Every frame is uploaded exactly once (the pair of T1:1059 is assembled on the device).
    ``k, v, prev`` continue an earlier part of the same clip (caches and the last device frame it returned).
    Returns (out_host, k_cache, v_cache, last_frame); out_host is complete when the function returns."""
    dev = torch.device(device) if device is not None else next(net.parameters()).device
    B, T, C, H, W = clip_host.shape
    up = 4 if getattr(net, "variant", "") == "super" else 1
    if out_host is None:
        out_host = torch.empty(B, T, getattr(net, "out_channels", C), H * up, W * up).pin_memory()
    main = torch.cuda.current_stream(dev)
    h2d, d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    slots = [torch.empty(B, C, H, W, device=dev) for _ in range(3)]         # frame j-1, j, j+1
    up_done = [torch.cuda.Event() for _ in range(3)]
    slot_free = [torch.cuda.Event() for _ in range(3)]

    def upload(j):
        s = j % 3
        with torch.cuda.stream(h2d):
            if j >= 3:
                h2d.wait_event(slot_free[s])                                  # frame j-3's last reader has run
            slots[s].copy_(clip_host[:, j], non_blocking=True)
            up_done[s].record(h2d)

    upload(0)
    pending = []
    for j in range(T):
        if j + 1 < T:
            upload(j + 1)
        main.wait_event(up_done[j % 3])
        cur = slots[j % 3]
        pre = (cur if prev is None else prev) if j == 0 else slots[(j - 1) % 3]
        o, k, v = net(torch.stack([pre, cur], dim=1), k, v)
        if j >= 1:
            slot_free[(j - 1) % 3].record(main)                               # frame j-1 is no longer read
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(d2h):
            d2h.wait_event(done)
            out_host[:, j].copy_(o, non_blocking=True)
            o.record_stream(d2h)
        pending.append(o)
        if len(pending) > 2:
            pending.pop(0)
    d2h.synchronize()
    return out_host, k, v, slots[(T - 1) % 3].clone()
