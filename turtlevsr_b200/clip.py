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
