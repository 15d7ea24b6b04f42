"""Tiled inference with device-resident per-tile histories.

Same contract as the reference's ``run_inference_patched`` (``basicsr/inference.py:172-246``): reflect-pad to a
multiple of 8, tiles of ``tile`` with stride ``tile - tile_overlap`` (last tile flush with the border), one
independent history per tile keyed ``"h-w"``, overlap-averaged and clamped to [0,1].  The reference moves every
tile's 10 cache tensors GPU->CPU->GPU each frame (INF:227-237); here the caches are views of HBM rings that stay on
the device, so passing the returned dicts back costs nothing.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


@torch.no_grad()
def run_inference_patched(img_lq_prev: torch.Tensor, img_lq_curr: torch.Tensor, model, device, tile: int,
                          tile_overlap: int, dataset_name: Optional[str] = None,
                          prev_patch_dict_k: Optional[Dict[str, list]] = None,
                          prev_patch_dict_v: Optional[Dict[str, list]] = None, img_multiple_of: int = 8, scale: int = 1,
                          model_type: str = "t0"):
    """img_lq_* : [B,C,H,W].  Returns (restored [B,C,H',W'] on ``device``, patch_dict_k, patch_dict_v)."""
    height, width = img_lq_curr.shape[2], img_lq_curr.shape[3]
    m = img_multiple_of
    Hp, Wp = ((height + m) // m) * m, ((width + m) // m) * m
    padh = Hp - height if height % m != 0 else 0
    padw = Wp - width if width % m != 0 else 0
    cur = F.pad(img_lq_curr.to(device), (0, padw, 0, padh), "reflect")
    prev = F.pad(img_lq_prev.to(device), (0, padw, 0, padh), "reflect")
    b, c, h, w = cur.shape
    tile = min(tile, h, w)
    assert tile % 8 == 0, "tile size should be multiple of 8"
    stride = tile - tile_overlap
    h_idx_list = list(range(0, h - tile, stride)) + [h - tile]
    w_idx_list = list(range(0, w - tile, stride)) + [w - tile]
    E = torch.zeros(b, c, h, w, device=device)
    Wt = torch.zeros_like(E)
    patch_dict_k, patch_dict_v = {}, {}
    for h_idx in h_idx_list:
        for w_idx in w_idx_list:
            p_cur = cur[..., h_idx:h_idx + tile, w_idx:w_idx + tile]
            p_prev = prev[..., h_idx:h_idx + tile, w_idx:w_idx + tile]
            if model_type == "SR":          # INF:213-219 (the SR arch upsamples x4 again internally)
                p_prev = F.interpolate(p_prev, scale_factor=1 / 4, mode="bicubic")
                p_cur = F.interpolate(p_cur, scale_factor=1 / 4, mode="bicubic")
            x = torch.stack((p_prev, p_cur), dim=1)
            key = f"{h_idx}-{w_idx}"
            old_k = old_v = None
            if prev_patch_dict_k is not None and prev_patch_dict_v is not None:
                old_k = [t if t is None or t.device == x.device else t.to(device) for t in prev_patch_dict_k[key]]
                old_v = [t if t is None or t.device == x.device else t.to(device) for t in prev_patch_dict_v[key]]
            out_patch, k_c, v_c = model(x.float(), old_k, old_v)
            patch_dict_k[key], patch_dict_v[key] = k_c, v_c           # ring views: stay in HBM
            E[..., h_idx:h_idx + tile, w_idx:w_idx + tile].add_(out_patch.to(device))
            Wt[..., h_idx:h_idx + tile, w_idx:w_idx + tile].add_(1.0)
    return torch.clamp(E.div_(Wt), 0, 1), patch_dict_k, patch_dict_v
