"""Shared host-side mirror of the reference arch interface.

The three reference arch files (turtle_arch.py "t0", turtle_t1_arch.py "t1",
turtlesuper_t1_arch.py "SR") expose ``make_model(opt)``, ``create_video_model(opt)`` and one
``nn.Module`` whose ``forward(x[B,2,C,H,W], k_cached, v_cached)`` returns
``(out[B,C,H,W], k_to_cache[8], v_to_cache[8])`` (T1:10-59, T1:1045-1132).  This module
re-creates that boundary:

* the module tree is a tree of *parameter holders* with exactly the reference's names,
  shapes and construction order (so ``torch.manual_seed(s); make_model(opt)`` draws the same
  random init as the reference, and ``load_state_dict(ref.state_dict(), strict=True)`` works --
  SURVEY.md Appendix B, 633 keys);
* ``forward`` hands the frame to :class:`turtlevsr_b200.engine.FrameEngine`, which runs the
  hand-written sm_100a kernels through the C-ABI (include/turtle_b200.h).  There is no torch
  fallback: without the CUDA library, or on CPU tensors, ``forward`` raises.
"""
from __future__ import annotations

import os

from typing import List, Optional

import torch
import torch.nn as nn

ATTN_ALIASES = {"MEST": "CHM", "CTS": "FHR"}   # names used by Turtle_Denoise_Davis.yml:42-55

_REQUIRED = [
    "n_colors", "dim", "Enc_blocks", "Middle_blocks", "Dec_blocks",
    "encoder1_attn_type1", "encoder1_attn_type2", "encoder2_attn_type1", "encoder2_attn_type2",
    "encoder3_attn_type1", "encoder3_attn_type2", "decoder1_attn_type1", "decoder1_attn_type2",
    "decoder2_attn_type1", "decoder2_attn_type2", "decoder3_attn_type1", "decoder3_attn_type2",
    "encoder1_ffw_type", "encoder2_ffw_type", "encoder3_ffw_type",
    "decoder1_ffw_type", "decoder2_ffw_type", "decoder3_ffw_type",
    "latent_attn_type1", "latent_attn_type2", "latent_attn_type3", "latent_ffw_type",
    "refinement_attn_type1", "refinement_attn_type2", "refinement_ffw_type", "use_both_input",
]


def model_kwargs_from_opt(opt: dict) -> dict:
    """Flat yml dict -> ctor kwargs, same defaults as the reference make_model (T1:10-53)."""
    for k in _REQUIRED:
        if k not in opt:
            raise KeyError(k)
    kw = {k: opt[k] for k in _REQUIRED if k != "n_colors"}
    kw.update(
        inp_channels=opt["n_colors"], out_channels=opt["n_colors"],
        num_refinement_blocks=opt.get("num_refinement_blocks", 1),
        ffn_expansion_factor=opt.get("ffn_expansion_factor", 1),
        bias=opt.get("bias", False),
        LayerNorm_type=opt.get("LayerNorm_type", "WithBias"),
        num_heads_blks=opt.get("num_heads_blks", [1, 2, 4, 8]),
        num_frames_tocache=opt.get("num_frames_tocache", 1),
        num_heads=opt.get("num_heads", [1, 1, 1, 1]),
    )
    return kw


# ----------------------------------------------------------------------------------------
# parameter holders (names / shapes / init order follow the reference classes)
# ----------------------------------------------------------------------------------------
class _NormBody(nn.Module):
    def __init__(self, dim, with_bias):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        if with_bias:
            self.bias = nn.Parameter(torch.zeros(dim))


class LayerNorm(nn.Module):                      # T1:102-112
    def __init__(self, dim, LayerNorm_type):
        super().__init__()
        self.body = _NormBody(dim, LayerNorm_type != "BiasFree")


class GatedFeedForward(nn.Module):               # T1:159-171
    def __init__(self, dim, ffn_expansion_factor, bias):
        super().__init__()
        hidden = int(dim * ffn_expansion_factor)
        self.project_in = nn.Conv2d(dim, hidden * 2, 1, bias=bias)
        self.dwconv = nn.Conv2d(hidden * 2, hidden * 2, 3, 1, 1, groups=hidden * 2, bias=bias)
        self.project_out = nn.Conv2d(hidden, dim, 1, bias=bias)


class FeedForward(nn.Module):                    # T1:181-202
    def __init__(self, c, FFN_Expand=2):
        super().__init__()
        self.conv4 = nn.Conv2d(c, FFN_Expand * c, 1, bias=True)
        self.conv5 = nn.Conv2d(FFN_Expand * c, c, 1, bias=True)
        self.gamma = nn.Parameter(torch.zeros((1, c, 1, 1)))


class ChannelAttention(nn.Module):               # T1:666-676
    def __init__(self, dim, num_heads, bias):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = nn.Conv2d(dim, dim * 3, 1, bias=bias)
        self.qkv_dwconv = nn.Conv2d(dim * 3, dim * 3, 3, 1, 1, groups=dim * 3, bias=bias)
        self.project_out = nn.Conv2d(dim, dim, 1, bias=bias)


class FrameHistoryRouter(ChannelAttention):      # T1:218-240
    def __init__(self, dim, num_heads, bias, num_frames_tocache=1):
        super().__init__(dim, num_heads, bias)
        self.num_frames_tocache = num_frames_tocache


class StateAlignBlock(nn.Module):                # T1:289-316
    def __init__(self, dim, num_heads, bias, num_frames_tocache, Scale_patchsize=1):
        super().__init__()
        self.num_heads = 1
        self.temperature = nn.Parameter(torch.ones(1, 1, 1))
        self.num_frames_tocache = num_frames_tocache
        ws = self.window_size = 2 * Scale_patchsize
        self.qk = nn.Conv2d(dim, dim * 2, 1, bias=bias)
        self.qk_dwconv = nn.Conv2d(dim * 2, dim * 2, 3, 1, 1, groups=dim * 2, bias=bias)
        self.v = nn.Conv2d(dim, dim, 1, bias=bias)
        self.v_dwconv = nn.Conv2d(dim, dim, 3, 1, 1, groups=dim, bias=bias)
        self.k2 = nn.Conv2d(dim, dim * 2, 1, bias=bias)
        self.k2_dwconv = nn.Conv2d(dim * 2, dim * 2, ws, ws, 1, groups=dim * 2, bias=bias)
        self.q2 = nn.Conv2d(dim, dim * 2, 1, bias=bias)
        self.q2_dwconv = nn.Conv2d(dim * 2, dim * 2, ws, ws, 1, groups=dim * 2, bias=bias)
        self.project_out = nn.Conv2d(dim, dim, 1, bias=bias)
        # the reference registers a None buffer, which adds no state-dict key (T1:313-316)
        self.register_buffer("local_mask", None)


class CausalHistoryModel(nn.Module):             # T1:612-625
    def __init__(self, dim, num_heads, bias, scale_patchsize, num_frames_tocache=1):
        super().__init__()
        self.spatial_aligner = StateAlignBlock(dim, num_heads, bias, num_frames_tocache,
                                               Scale_patchsize=scale_patchsize)
        self.ChanAttn = FrameHistoryRouter(dim, num_heads, bias)
        self.kv = nn.Conv2d(dim, dim * 2, 1, bias=bias)
        self.kv_dwconv = nn.Conv2d(dim * 2, dim * 2, 3, 1, 1, groups=dim * 2, bias=bias)
        self.num_heads = num_heads


class ReducedAttn(nn.Module):                    # T1:704-734
    def __init__(self, c, DW_Expand=2.0):
        super().__init__()
        dw = int(c * DW_Expand)
        self.conv1 = nn.Conv2d(c, dw, 1, bias=True)
        self.conv2 = nn.Conv2d(dw, dw, 3, 1, 1, groups=dw, bias=True)
        self.conv3 = nn.Conv2d(dw, c, 1, bias=True)
        self.beta = nn.Parameter(torch.zeros((1, c, 1, 1)))


class TurtleAttnBlock(nn.Module):                # T1:746-802
    def __init__(self, dim, ffn_expansion_factor, bias, LayerNorm_type, num_heads=1, Scale_patchsize=1,
                 attention_type="channel", FFW_type="GFFW", num_frames_tocache=1):
        super().__init__()
        attention_type = ATTN_ALIASES.get(attention_type, attention_type)
        self.attention_type, self.FFW_type = attention_type, FFW_type
        self.norm1 = LayerNorm(dim, LayerNorm_type)
        if attention_type == "Channel":
            self.attn = ChannelAttention(dim, num_heads, bias)
        elif attention_type == "ReducedAttn":
            self.attn = ReducedAttn(dim)
        elif attention_type == "FHR":
            self.attn = FrameHistoryRouter(dim, num_heads, bias, num_frames_tocache)
        elif attention_type == "CHM":
            self.attn = CausalHistoryModel(dim, num_heads, bias, Scale_patchsize, num_frames_tocache)
        elif attention_type == "NoAttn":
            self.attn = None
        else:                                    # same failure mode as T1:790-792
            print(attention_type, " Not defined")
            raise SystemExit()
        self.norm2 = LayerNorm(dim, LayerNorm_type)
        if FFW_type == "GFFW":
            self.ffn = GatedFeedForward(dim, ffn_expansion_factor, bias)
        elif FFW_type == "FFW":
            self.ffn = FeedForward(dim)
        else:
            print(FFW_type, " Not defined")
            raise SystemExit()


class LevelBlock(nn.Module):                     # T1:813-854
    def __init__(self, dim, ffn_expansion_factor, bias, LayerNorm_type, num_blocks, attn_type1="Channel",
                 attn_type2="CHM", FFW_type="GFFW", num_frames_tocache=1, num_heads=1, Scale_patchsize=1):
        super().__init__()
        self.num_blocks = num_blocks
        self.dim, self.num_heads = dim, num_heads
        self.num_frames_tocache, self.Scale_patchsize = num_frames_tocache, Scale_patchsize
        types = [attn_type1] * (num_blocks - 1) + [attn_type2]
        self.transformer_blocks = nn.ModuleList([
            TurtleAttnBlock(dim=dim, num_heads=num_heads, ffn_expansion_factor=ffn_expansion_factor, bias=bias,
                            LayerNorm_type=LayerNorm_type, attention_type=t, FFW_type=FFW_type,
                            num_frames_tocache=num_frames_tocache, Scale_patchsize=Scale_patchsize)
            for t in types])


class LatentCacheBlock(nn.Module):               # T1:867-917
    def __init__(self, dim, ffn_expansion_factor, bias, LayerNorm_type, num_blocks, attn_type1="FHR",
                 attn_type2="Channel", attn_type3="FHR", FFW_type="GFFW", num_frames_tocache=1, num_heads=1):
        super().__init__()
        self.num_blocks = num_blocks
        self.dim, self.num_heads, self.num_frames_tocache = dim, num_heads, num_frames_tocache
        self.Scale_patchsize = 1
        if num_blocks < 2:
            print("LatentCacheBlock should have more than 2 layers")
            raise SystemExit()
        types = [attn_type1] + [attn_type2] * (num_blocks - 2) + [attn_type3]
        self.transformer_blocks = nn.ModuleList([
            TurtleAttnBlock(dim=dim, num_heads=num_heads, ffn_expansion_factor=ffn_expansion_factor, bias=bias,
                            LayerNorm_type=LayerNorm_type, attention_type=t, FFW_type=FFW_type,
                            num_frames_tocache=num_frames_tocache)
            for t in types])


class _Resample(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        # body.1 is the parameter-free Pixel(Un)Shuffle in the reference (T1:136-154)
        self.body = nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1, bias=False))


class TurtleNet(nn.Module):
    """Drop-in for Turtle / Turtle_t1 / TurtleSuper_t1 (selected by ``variant``)."""

    variant = "t1"

    def __init__(self, inp_channels, out_channels, dim, Enc_blocks, Middle_blocks, Dec_blocks, num_heads,
                 num_refinement_blocks, ffn_expansion_factor, bias, LayerNorm_type, num_heads_blks,
                 encoder1_attn_type1, encoder1_attn_type2, encoder2_attn_type1, encoder2_attn_type2,
                 encoder3_attn_type1, encoder3_attn_type2, decoder1_attn_type1, decoder1_attn_type2,
                 decoder2_attn_type1, decoder2_attn_type2, decoder3_attn_type1, decoder3_attn_type2,
                 encoder1_ffw_type, encoder2_ffw_type, encoder3_ffw_type,
                 decoder1_ffw_type, decoder2_ffw_type, decoder3_ffw_type,
                 latent_attn_type1, latent_attn_type2, latent_attn_type3, latent_ffw_type,
                 refinement_attn_type1, refinement_attn_type2, refinement_ffw_type,
                 use_both_input, num_frames_tocache):
        super().__init__()
        self.out_channels = out_channels
        if use_both_input:
            inp_channels *= 2
        self.use_both_input = use_both_input
        self.num_heads = num_heads
        self.dim = dim
        self.ffn_expansion_factor = ffn_expansion_factor
        common = dict(ffn_expansion_factor=ffn_expansion_factor, bias=bias, LayerNorm_type=LayerNorm_type)
        K = num_frames_tocache
        # construction order == reference order (T1:975-1041): identical RNG draw sequence
        self.input_projection = nn.Conv2d(inp_channels, dim, 3, 1, 1, bias=bias)
        self.encoder_level1 = LevelBlock(dim=dim, num_blocks=Enc_blocks[0], attn_type1=encoder1_attn_type1,
                                         attn_type2=encoder1_attn_type2, FFW_type=encoder1_ffw_type,
                                         num_frames_tocache=K, num_heads=num_heads[0], **common)
        self.down1_2 = _Resample(dim, dim // 2)
        self.encoder_level2 = LevelBlock(dim=dim * 2, num_blocks=Enc_blocks[1], attn_type1=encoder2_attn_type1,
                                         attn_type2=encoder2_attn_type2, FFW_type=encoder2_ffw_type,
                                         num_frames_tocache=K, num_heads=num_heads[1], **common)
        self.down2_3 = _Resample(dim * 2, dim)
        self.encoder_level3 = LevelBlock(dim=dim * 4, num_blocks=Enc_blocks[2], attn_type1=encoder3_attn_type1,
                                         attn_type2=encoder3_attn_type2, FFW_type=encoder3_ffw_type,
                                         num_frames_tocache=K, num_heads=num_heads[2], **common)
        self.down3_4 = _Resample(dim * 4, dim * 2)
        self.latent = LatentCacheBlock(dim=dim * 8, num_blocks=Middle_blocks, attn_type1=latent_attn_type1,
                                       attn_type2=latent_attn_type2, attn_type3=latent_attn_type3,
                                       FFW_type=latent_ffw_type, num_frames_tocache=K, num_heads=num_heads[3],
                                       **common)
        self.up4_3 = _Resample(dim * 8, dim * 16)
        self.reduce_chan_level3 = nn.Conv2d(dim * 8, dim * 4, 1, bias=bias)
        self.decoder_level3 = LevelBlock(dim=dim * 4, num_blocks=Dec_blocks[0], attn_type1=decoder1_attn_type1,
                                         attn_type2=decoder1_attn_type2, FFW_type=decoder1_ffw_type,
                                         num_frames_tocache=K, num_heads=num_heads[2], Scale_patchsize=2, **common)
        self.up3_2 = _Resample(dim * 4, dim * 8)
        self.reduce_chan_level2 = nn.Conv2d(dim * 4, dim * 2, 1, bias=bias)
        self.decoder_level2 = LevelBlock(dim=dim * 2, num_blocks=Dec_blocks[1], attn_type1=decoder2_attn_type1,
                                         attn_type2=decoder2_attn_type2, FFW_type=decoder2_ffw_type,
                                         num_frames_tocache=K, num_heads=num_heads[1], Scale_patchsize=4, **common)
        self.up2_1 = _Resample(dim * 2, dim * 4)
        self.reduce_chan_level1 = nn.Conv2d(dim * 2, dim, 1, bias=bias)
        self.decoder_level1 = LevelBlock(dim=dim, num_blocks=Dec_blocks[2], attn_type1=decoder3_attn_type1,
                                         attn_type2=decoder3_attn_type2, FFW_type=decoder3_ffw_type,
                                         num_frames_tocache=2,           # hard-coded in the reference, T1:1027
                                         num_heads=num_heads[0], Scale_patchsize=8, **common)
        self.refinement = LevelBlock(dim=dim, num_blocks=num_refinement_blocks,
                                     attn_type1=refinement_attn_type1, attn_type2=refinement_attn_type2,
                                     FFW_type=refinement_ffw_type, num_frames_tocache=K, num_heads=num_heads[0],
                                     **common)
        self.ending = nn.Conv2d(dim, out_channels, 3, 1, 1, bias=True)
        self.padder_size = (2 ** 3) * 4
        # execution state (not part of the state dict)
        self.precision = "fp32"        # "fp32": CUDA-core fp32 everywhere; "tf32": tcgen05 TF32 contractions
        self.cuda_graphs = False
        self.half_intermediates = True  # tf32 mode: FFN-side intermediates stored fp16 (same 10-bit mantissa as TF32)
        self.fuse_gffw = False          # tf32 mode: GatedFeedForward as the single kernel of csrc/gffw_fused.cu (parity-equal
                                        # to the three-kernel schedule, measured slower on B200: DESIGN.md section 3)
        self.gffw_tail = False          # tf32 mode: depthwise + gate + project_out of GatedFeedForward as one kernel
                                        # (turtle_gffw_tail; parity-equal, measured slower: DESIGN.md section 3)
        # tf32 mode: StateAlignBlock aggregation on the tensor cores (csrc/sab_agg_tc.cu: dense key-box contraction + far
        # top-k gather over an fp16 copy of the value rows); False / TURTLE_SAB_AGG_TC=0: the CUDA-core quad kernel
        self.sab_agg_tc = os.environ.get("TURTLE_SAB_AGG_TC", "1") != "0"
        self._engine = None

    # -- public knobs ------------------------------------------------------------------
    def set_precision(self, mode: str) -> "TurtleNet":
        """``"fp32"`` = exact mode (bit-exact top-k contract); ``"tf32"`` = tensor-core mode."""
        if mode not in ("fp32", "tf32"):
            raise ValueError(mode)
        self.precision = mode
        return self

    def enable_cuda_graphs(self, flag: bool = True, max_graphs: int = 0) -> "TurtleNet":
        """Replay steady-state frames (all history rings full, caches passed back unchanged) from CUDA graphs, one per
        joint ring state (history.RING_PERIOD of them per clip); other frames run eagerly.  Results are bit-identical to
        the eager path -- the same kernels in the same order."""
        self.cuda_graphs = bool(flag)
        if max_graphs:
            # each history (clip, or tile of the tile-by-tile loop) needs RING_PERIOD + 1 graphs; a cached graph pins its
            # history's rings in HBM until it is evicted (oldest first), so keep this near 6 x the histories in flight
            self.cuda_graph_limit = int(max_graphs)
            if self._engine is not None:
                self._engine.max_graphs = int(max_graphs)
        return self

    def invalidate_packed_weights(self) -> None:
        """Call after mutating parameters in place (load_state_dict does this automatically)."""
        if self._engine is not None:
            self._engine.invalidate()

    def load_state_dict(self, state_dict, *args, **kwargs):
        r = super().load_state_dict(state_dict, *args, **kwargs)
        self.invalidate_packed_weights()
        return r

    def _apply(self, fn, *args, **kwargs):
        r = super()._apply(fn, *args, **kwargs)
        self.invalidate_packed_weights()
        return r

    # -- the drop-in forward -------------------------------------------------------------
    def forward(self, inp_img_, k_cached: Optional[List] = None, v_cached: Optional[List] = None):
        from ..engine import FrameEngine
        if self.training and torch.is_grad_enabled():
            # training step (cfg 5, VRM:78-108): the differentiable graph of training.py (library autograd kernels);
            # the hand-written inference kernels have no backward yet
            from ..training import autograd_forward
            return autograd_forward(self, inp_img_, k_cached, v_cached)
        if self._engine is None:
            self._engine = FrameEngine(self)
        return self._engine.forward(inp_img_, k_cached, v_cached)

    def check_image_size(self, x):               # T1:1134-1140 (host helper, kept for API parity)
        h, w = x.shape[-2:]
        ph = (self.padder_size - h % self.padder_size) % self.padder_size
        pw = (self.padder_size - w % self.padder_size) % self.padder_size
        return torch.nn.functional.pad(x, (0, pw, 0, ph))
