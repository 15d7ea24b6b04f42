"""Tiled inference with device-resident per-tile histories.

Same contract as the reference's ``run_inference_patched`` (``basicsr/inference.py:172-246``): reflect-pad to a
multiple of 8, tiles of ``tile`` with stride ``tile - tile_overlap`` (last tile flush with the border), one
independent history per tile keyed ``"h-w"``, overlap-averaged and clamped to [0,1].  The reference moves every
tile's 10 cache tensors GPU->CPU->GPU each frame (INF:227-237); here the caches are views of HBM rings that stay on
the device, so passing the returned dicts back costs nothing.

``batch_tiles=True`` (this package's addition): all tiles of a frame have the same size, so they
are stacked into ONE batch and restored by a single forward (B = number of tiles) whose history rings carry the tile
index as their batch dimension; the per-frame cost is then one set of ~520 large launches instead of one set of small,
launch-bound ones per tile.  The batched caches travel under the key ``BATCH_KEY``; per-tile keys still map to
(read-only) slices so that code inspecting the dicts keeps working.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

BATCH_KEY = "__all_tiles__"


@torch.no_grad()
def run_inference_patched(img_lq_prev: torch.Tensor, img_lq_curr: torch.Tensor, model, device, tile: int,
                          tile_overlap: int, dataset_name: Optional[str] = None,
                          prev_patch_dict_k: Optional[Dict[str, list]] = None,
                          prev_patch_dict_v: Optional[Dict[str, list]] = None, img_multiple_of: int = 8, scale: int = 1,
                          model_type: str = "t0", batch_tiles: bool = False):
    """img_lq_* : [B,C,H,W].  Returns (restored [B,C,H',W'] on ``device``, patch_dict_k, patch_dict_v)."""
    height, width = img_lq_curr.shape[2], img_lq_curr.shape[3]
    m = img_multiple_of
    Hp, Wp = ((height + m) // m) * m, ((width + m) // m) * m
    padh = Hp - height if height % m != 0 else 0
    padw = Wp - width if width % m != 0 else 0
    b, c = img_lq_curr.shape[:2]
    h, w = height + padh, width + padw
    tile = min(tile, h, w)
    assert tile % 8 == 0, "tile size should be multiple of 8"
    stride = tile - tile_overlap
    h_idx_list = list(range(0, h - tile, stride)) + [h - tile]
    w_idx_list = list(range(0, w - tile, stride)) + [w - tile]
    patch_dict_k, patch_dict_v = {}, {}
    cur = img_lq_curr.to(device)
    if batch_tiles and cur.is_cuda and model_type != "SR" and b == 1 and hasattr(model, "set_precision"):
        # fused path: the reflect padding, the tile cut and the (previous, current) pairing are ONE gather kernel reading
        # the un-padded frames, and the overlap average + clamp is ONE gather over the restored tiles (no E / W
        # accumulators) -- turtle_tile_gather / turtle_tile_blend, csrc/frameio.cu
        import ctypes as C_
        from .capi import call
        nt = len(h_idx_list) * len(w_idx_list)
        ys = (C_.c_int32 * len(h_idx_list))(*h_idx_list)
        xs = (C_.c_int32 * len(w_idx_list))(*w_idx_list)
        src_c = cur.float().contiguous()
        src_p = img_lq_prev.to(device).float().contiguous()
        x = torch.empty(nt, 2, c, tile, tile, device=cur.device)
        stream = torch.cuda.current_stream(cur.device).cuda_stream
        with torch.cuda.device(cur.device):
            call("turtle_tile_gather", src_p.data_ptr(), src_c.data_ptr(), x.data_ptr(), c, height, width, tile, ys,
                 len(h_idx_list), xs, len(w_idx_list), stream)
            old_k = old_v = None
            if prev_patch_dict_k is not None and prev_patch_dict_v is not None:
                old_k, old_v = prev_patch_dict_k[BATCH_KEY], prev_patch_dict_v[BATCH_KEY]
            out, k_c, v_c = model(x, old_k, old_v)
            res = torch.empty(1, out.shape[1], h, w, device=cur.device)
            call("turtle_tile_blend", out.contiguous().data_ptr(), res.data_ptr(), out.shape[1], h, w, tile, ys,
                 len(h_idx_list), xs, len(w_idx_list), 1, stream)
        patch_dict_k[BATCH_KEY], patch_dict_v[BATCH_KEY] = k_c, v_c
        i = 0
        for hi in h_idx_list:
            for wi in w_idx_list:
                patch_dict_k[f"{hi}-{wi}"] = [None if t is None else t[i:i + 1] for t in k_c]
                patch_dict_v[f"{hi}-{wi}"] = [None if t is None else t[i:i + 1] for t in v_c]
                i += 1
        return res, patch_dict_k, patch_dict_v
    cur = F.pad(cur, (0, padw, 0, padh), "reflect")
    prev = F.pad(img_lq_prev.to(device), (0, padw, 0, padh), "reflect")
    E = torch.zeros(b, c, h, w, device=device)
    Wt = torch.zeros_like(E)
    if batch_tiles:
        if b != 1:
            raise ValueError("batch_tiles stacks the tiles along the batch axis: frames must come one at a time (B=1)")
        pos = [(hi, wi) for hi in h_idx_list for wi in w_idx_list]
        p_cur = torch.cat([cur[..., hi:hi + tile, wi:wi + tile] for hi, wi in pos], 0)
        p_prev = torch.cat([prev[..., hi:hi + tile, wi:wi + tile] for hi, wi in pos], 0)
        if model_type == "SR":
            p_prev = F.interpolate(p_prev, scale_factor=1 / 4, mode="bicubic")
            p_cur = F.interpolate(p_cur, scale_factor=1 / 4, mode="bicubic")
        old_k = old_v = None
        if prev_patch_dict_k is not None and prev_patch_dict_v is not None:
            old_k, old_v = prev_patch_dict_k[BATCH_KEY], prev_patch_dict_v[BATCH_KEY]
        out, k_c, v_c = model(torch.stack((p_prev, p_cur), dim=1).float(), old_k, old_v)
        patch_dict_k[BATCH_KEY], patch_dict_v[BATCH_KEY] = k_c, v_c
        for i, (hi, wi) in enumerate(pos):
            E[..., hi:hi + tile, wi:wi + tile].add_(out[i:i + 1])
            Wt[..., hi:hi + tile, wi:wi + tile].add_(1.0)
            patch_dict_k[f"{hi}-{wi}"] = [None if t is None else t[i:i + 1] for t in k_c]
            patch_dict_v[f"{hi}-{wi}"] = [None if t is None else t[i:i + 1] for t in v_c]
        return torch.clamp(E.div_(Wt), 0, 1), patch_dict_k, patch_dict_v
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
