"""Training step of BASELINE config 5: the reference's ``VideoRestorationModel.optimize_parameters``
(VRM:78-108) as a data-parallel step -- one process per GPU, a bucketed NCCL gradient all-reduce that overlaps the
backward pass, and the optimizer as two passes over ONE flat fp32 buffer on hand-written kernels.

What runs where (stated plainly):

* forward + backward: ``autograd_forward`` below -- the same arch tree evaluated with differentiable torch ops
  (cuDNN / cuBLAS / ATen library kernels on the GPU, except the channel LayerNorm and the depthwise 3x3 convs, whose
  forward and backward are hand-written: csrc/ln2d_train.cu, csrc/dwconv_train.cu; BPTT through the history caches exactly as the reference does it,
  the caches stay attached to the graph across the frames of a clip).  The hand-written inference kernels have no
  backward yet, so ``TurtleNet.forward`` routes here when ``training and grad enabled``;
* gradient exchange: ``GradBuckets`` -- every ``.grad`` is a view of one flat buffer cut into ~32 MB buckets in reverse
  construction order; a post-accumulate hook fires an async ``all_reduce`` on a bucket as soon as its last gradient is
  written (NCCL over NVLink on the GPU box, gloo in the CPU tests).  59,079,548 fp32 gradients = 236 MB per step;
* unscale + non-finite check + AdamW: ``turtle_grad_check_finite`` and ``turtle_adamw_flat`` (csrc/train_opt.cu) over
  the flat parameter / gradient / moment buffers: 7 x 4 B per parameter = 1.65 GB of HBM traffic per step instead of
  ~633 small per-tensor launches.  The 1/world_size of the gradient mean and the 1/loss_scale of the scaler are one
  factor applied inside the AdamW kernel.  No CUDA library => ``FlatAdamW`` raises (no CPU optimizer fallback).
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------
# differentiable frame forward over the parameter-holder tree (archs/_common.py)
# ----------------------------------------------------------------------------------------
_LN_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


class _ChannelLayerNorm(torch.autograd.Function):
    """WithBias_LayerNorm (T1:83-112) on the sm_100a kernels turtle_ln2d_fwd / turtle_ln2d_bwd: one launch forward, two
    backward, instead of the ~30 ATen launches autograd records for the mean / var / sqrt / div / mul / add chain."""

    @staticmethod
    def forward(ctx, x, w, b, out_dtype=torch.float32):
        """``out_dtype``: fp32 as autocast leaves the reference's formula, or the autocast dtype when the only consumer
        is a convolution (which would cast it there: the same single rounding, without the cast launches)."""
        from . import capi
        x = x.contiguous()
        B, C, H, W = x.shape
        y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
        mean = torch.empty(B * H * W, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        wf, bf = w.detach().float().contiguous(), b.detach().float().contiguous()
        capi.call("turtle_ln2d_fwd_cast", x.data_ptr(), _LN_DTYPES[x.dtype], wf.data_ptr(), bf.data_ptr(), y.data_ptr(),
                  _LN_DTYPES[out_dtype], mean.data_ptr(), rstd.data_ptr(), B, C, H * W,
                  torch.cuda.current_stream(x.device).cuda_stream)
        ctx.save_for_backward(x, wf, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import capi
        x, wf, mean, rstd = ctx.saved_tensors
        B, C, H, W = x.shape
        dy = dy.contiguous()
        if dy.dtype not in _LN_DTYPES:
            dy = dy.float()
        dx = torch.empty_like(x)
        dw = torch.empty(C, device=x.device, dtype=torch.float32)
        db = torch.empty_like(dw)
        nbytes = capi.load().turtle_ln2d_bwd_workspace(C, B * H * W)
        wsp = torch.empty(nbytes // 4, device=x.device, dtype=torch.float32)
        capi.call("turtle_ln2d_bwd_cast", dy.data_ptr(), _LN_DTYPES[dy.dtype], x.data_ptr(), _LN_DTYPES[x.dtype],
                  wf.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dx.data_ptr(), dw.data_ptr(), db.data_ptr(),
                  wsp.data_ptr(), B, C, H * W, torch.cuda.current_stream(x.device).cuda_stream)
        return dx, dw, db, None


class _Depthwise3x3(torch.autograd.Function):
    """nn.Conv2d(C, C, 3, 1, 1, groups=C) on turtle_dwconv3x3_nchw / _wgrad: forward, input gradient and weight / bias
    gradient as one launch each (+ a tiny fixed-order reduction) instead of ATen's three depthwise kernels."""

    @staticmethod
    def forward(ctx, x, w, b):
        from . import capi
        x = x.contiguous()
        B, C, H, W = x.shape
        w9 = w.detach().float().reshape(C, 9).contiguous()
        bf = None if b is None else b.detach().float().contiguous()
        y = torch.empty_like(x)
        capi.call("turtle_dwconv3x3_nchw", x.data_ptr(), _LN_DTYPES[x.dtype], w9.data_ptr(),
                  None if bf is None else bf.data_ptr(), y.data_ptr(), B, C, H, W, 0,
                  torch.cuda.current_stream(x.device).cuda_stream)
        ctx.save_for_backward(x, w9)
        ctx.has_bias = b is not None
        ctx.w_shape = w.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import capi
        x, w9 = ctx.saved_tensors
        B, C, H, W = x.shape
        dy = dy.to(x.dtype).contiguous()
        st = torch.cuda.current_stream(x.device).cuda_stream
        dx = torch.empty_like(x)
        capi.call("turtle_dwconv3x3_nchw", dy.data_ptr(), _LN_DTYPES[x.dtype], w9.data_ptr(), None, dx.data_ptr(),
                  B, C, H, W, 1, st)
        dw9 = torch.empty(C, 9, device=x.device, dtype=torch.float32)
        db = torch.empty(C, device=x.device, dtype=torch.float32) if ctx.has_bias else None
        nbytes = capi.load().turtle_dwconv3x3_nchw_wgrad_workspace(B, C, H, W)
        wsp = torch.empty(nbytes // 4, device=x.device, dtype=torch.float32)
        capi.call("turtle_dwconv3x3_nchw_wgrad", x.data_ptr(), dy.data_ptr(), _LN_DTYPES[x.dtype], dw9.data_ptr(),
                  None if db is None else db.data_ptr(), wsp.data_ptr(), B, C, H, W, st)
        return dx, dw9.view(ctx.w_shape), db


def _aligned(t: Tensor, nbytes: int = 32) -> Tensor:
    """A dense tensor whose storage starts on the kernels' vector alignment (an offset view of a larger buffer may not)."""
    t = t.contiguous()
    return t if t.data_ptr() % nbytes == 0 else t.clone()


class _GeluGate(torch.autograd.Function):
    """`a, g = u.chunk(2, dim=1); F.gelu(a) * g` (GatedFeedForward, T1:175-176) on turtle_gelu_gate_nchw / _bwd: one launch
    forward and one backward instead of ATen's strided elementwise kernels on the channel-chunk views (gelu, mul, two
    mul-backwards, gelu_backward and the zero-filled cat of the halves)."""

    @staticmethod
    def forward(ctx, u):
        from . import capi
        u = _aligned(u)
        B, C2, H, W = u.shape
        y = torch.empty(B, C2 // 2, H, W, device=u.device, dtype=u.dtype)
        capi.call("turtle_gelu_gate_nchw", u.data_ptr(), _LN_DTYPES[u.dtype], y.data_ptr(), B, C2 // 2, H * W,
                  torch.cuda.current_stream(u.device).cuda_stream)
        ctx.save_for_backward(u)
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import capi
        (u,) = ctx.saved_tensors
        B, C2, H, W = u.shape
        if dy.dtype != u.dtype:
            dy = dy.to(u.dtype)
        dy = _aligned(dy)
        du = torch.empty_like(u)
        capi.call("turtle_gelu_gate_nchw_bwd", u.data_ptr(), dy.data_ptr(), _LN_DTYPES[u.dtype], du.data_ptr(), B, C2 // 2,
                  H * W, torch.cuda.current_stream(u.device).cuda_stream)
        return du


def _gelu_gate(u: Tensor) -> Tensor:
    """gelu(first half of the channels) * second half."""
    if (u.is_cuda and u.dtype in _LN_DTYPES and u.shape[1] % 2 == 0 and (u.shape[2] * u.shape[3]) % 8 == 0
            and os.environ.get("TURTLE_TRAIN_GATE", "1") != "0"):
        return _GeluGate.apply(u)
    a, g = u.chunk(2, dim=1)
    return F.gelu(a) * g


def _dw(conv, x: Tensor) -> Tensor:
    """A depthwise conv module of the arch tree: the hand-written kernels for the 3x3 / stride 1 / pad 1 case on CUDA."""
    if (x.is_cuda and x.dtype in _LN_DTYPES and conv.kernel_size == (3, 3) and conv.stride == (1, 1)
            and conv.padding == (1, 1) and conv.groups == conv.in_channels == conv.out_channels
            and x.shape[0] * x.shape[1] <= 65535):
        if torch.is_autocast_enabled("cuda"):                    # what autocast would do to conv2d's input
            x = x.to(torch.get_autocast_dtype("cuda"))
        return _Depthwise3x3.apply(x, conv.weight, conv.bias)
    return conv(x)


def _layernorm(norm, x: Tensor, conv_next: bool = False) -> Tensor:                      # T1:83-112
    """``conv_next``: the result feeds convolutions only, so under autocast it may leave the kernel in the autocast dtype."""
    b = getattr(norm.body, "bias", None)
    if x.is_cuda and b is not None and x.dtype in _LN_DTYPES:
        od = torch.float32
        if (conv_next and torch.is_autocast_enabled("cuda") and os.environ.get("TURTLE_TRAIN_LN_CAST", "1") != "0"
                and torch.get_autocast_dtype("cuda") in _LN_DTYPES):
            od = torch.get_autocast_dtype("cuda")
        return _ChannelLayerNorm.apply(x, norm.body.weight, b, od)
    # BiasFree variant (unused by the shipped ymls) and the CPU host-logic tests: the reference's formula on torch ops
    w = norm.body.weight.view(1, -1, 1, 1)
    mu = x.mean(dim=1, keepdim=True)
    var = (x - mu).pow(2).mean(dim=1, keepdim=True)
    if b is None:
        return x / torch.sqrt(var + 1e-5) * w
    return (x - mu) / torch.sqrt(var + 1e-5) * w + b.view(1, -1, 1, 1)


class _UnitRows(torch.autograd.Function):
    """F.normalize(x, dim=-1) on turtle_rownorm_fwd / _bwd for rows that are whole (image, channel) planes of a channel
    chunk of an NCHW map: one launch forward (reading the 16-bit view directly; fp32 out as autocast's normalize) and one
    backward instead of the strided cast + norm / clamp / div chain and its broadcast-heavy autograd backward."""

    @staticmethod
    def forward(ctx, x):
        from . import capi
        b, hd, cc, n = x.shape
        y = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        denom = torch.empty(b * hd * cc, device=x.device, dtype=torch.float32)
        capi.call("turtle_rownorm_fwd", x.data_ptr(), _LN_DTYPES[x.dtype], x.stride(0), b, hd * cc, n, y.data_ptr(),
                  denom.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream)
        ctx.save_for_backward(y, denom)
        ctx.x_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        from . import capi
        y, denom = ctx.saved_tensors
        dy = _aligned(dy.float())
        dx = torch.empty(y.shape, device=y.device, dtype=ctx.x_dtype)
        capi.call("turtle_rownorm_bwd", dy.data_ptr(), y.data_ptr(), denom.data_ptr(), _LN_DTYPES[ctx.x_dtype],
                  y.numel() // y.shape[-1], y.shape[-1], dx.data_ptr(), torch.cuda.current_stream(y.device).cuda_stream)
        return dx


def _unit_rows(x: Tensor) -> Tensor:
    """F.normalize(x, dim=-1) (T1:686-687); the fused kernels when the rows are the planes of a channel chunk."""
    if (x.is_cuda and x.dim() == 4 and x.dtype in _LN_DTYPES and x.shape[-1] % 4 == 0 and x.stride(3) == 1
            and x.stride(2) == x.shape[3] and x.stride(1) == x.shape[2] * x.shape[3]
            and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and os.environ.get("TURTLE_TRAIN_ROWNORM", "1") != "0"):
        return _UnitRows.apply(x)
    return F.normalize(x, dim=-1)


def _patches(t: Tensor, ws: int) -> Tensor:
    """'b d (p1 h) (p2 w) -> b (h w) (p1 p2 d)' (T1:573)."""
    b, d, h, w = t.shape
    gh, gw = h // ws, w // ws
    return t.reshape(b, d, ws, gh, ws, gw).permute(0, 3, 5, 2, 4, 1).reshape(b, gh * gw, ws * ws * d)


def _unpatch(o: Tensor, ws: int, d: int, h: int, w: int) -> Tensor:
    """'b f (h w) (p1 p2 d) -> b f d (p1 h) (p2 w)' (T1:602-604)."""
    b, f = o.shape[:2]
    gh, gw = h // ws, w // ws
    return o.reshape(b, f, gh, gw, ws, ws, d).permute(0, 1, 6, 4, 2, 5, 3).reshape(b, f, d, h, w)


_POSENC_CACHE: dict = {}


def _posenc(c: int, h: int, w: int, device, dtype) -> Tensor:   # T0:412-439
    key = (c, h, w, str(device), dtype)
    if key not in _POSENC_CACHE:        # built on the host once (also keeps the H2D copy out of CUDA-graph captures)
        _POSENC_CACHE[key] = _posenc_build(c, h, w, device, dtype)
    return _POSENC_CACHE[key]


def _posenc_build(c: int, h: int, w: int, device, dtype) -> Tensor:
    if c % 4 != 0:
        raise ValueError("Cannot use sin/cos positional encoding with odd dimension (got dim={:d})".format(c))
    half = c // 2
    div = torch.exp(torch.arange(0., half, 2) * -(math.log(10000.0) / half))
    pw = (torch.arange(0., w).unsqueeze(1) * div).t()            # [half/2, w]
    ph = (torch.arange(0., h).unsqueeze(1) * div).t()
    pe = torch.zeros(c, h, w)
    pe[0:half:2] = torch.sin(pw)[:, None, :]
    pe[1:half:2] = torch.cos(pw)[:, None, :]
    pe[half::2] = torch.sin(ph)[:, :, None]
    pe[half + 1::2] = torch.cos(ph)[:, :, None]
    return pe.to(device=device, dtype=dtype)


def _local_mask(gh: int, gw: int, device, dtype, radius: int = 4) -> Tensor:   # T1:448-464, on the device
    yy, xx = torch.meshgrid(torch.arange(gh, device=device), torch.arange(gw, device=device), indexing="ij")
    yy, xx = yy.reshape(-1), xx.reshape(-1)
    return ((yy[:, None] - yy[None]).abs() + (xx[:, None] - xx[None]).abs() <= radius).to(dtype)


def _clipped_softmax(z: Tensor) -> Tensor:                      # T1:115-132
    dead = z == 0
    p = torch.softmax(z.masked_fill(dead, float("-inf")), dim=-1).masked_fill(dead, 0)
    return p / p.sum(dim=-1, keepdim=True)


def _gated_ffw(m, x):                                           # T1:173-178
    return m.project_out(_gelu_gate(_dw(m.dwconv, m.project_in(x))))


def _plain_ffw(m, x):                                           # T1:204-210
    return m.conv5(F.gelu(m.conv4(x))) * m.gamma


def _reduced_attn(m, x):                                        # T1:736-742
    return m.conv3(F.gelu(_dw(m.conv2, m.conv1(x)))) * m.beta


def _channel_attn(m, x, k_hist=None, v_hist=None, keep_frames=None):   # T1:680-702, 243-286
    b, c, h, w = x.shape
    hd = m.num_heads
    q, k, v = (t.reshape(b, hd, c // hd, h * w) for t in _dw(m.qkv_dwconv, m.qkv(x)).chunk(3, dim=1))
    q, k = _unit_rows(q), _unit_rows(k)
    if k_hist is not None and v_hist is not None:
        k, v = torch.cat([k_hist, k], dim=2), torch.cat([v_hist, v], dim=2)
    attn = torch.softmax((q @ k.transpose(-1, -2)) * m.temperature, dim=-1)
    out = m.project_out((attn @ v).reshape(b, c, h, w))
    if keep_frames is None:
        return out, None, None
    keep = int(keep_frames * c / hd)
    return out, k[:, :, -keep:], v[:, :, -keep:]


def _state_align(m, x, variant, k_hist, v_hist):                # T1:548-610 / T0:459-533
    b, c, h, w = x.shape
    ws, keep = m.window_size, m.num_frames_tocache
    t0 = variant == "t0"
    x_qk = x + _posenc(c, h, w, x.device, x.dtype) if t0 else x
    q, k = _dw(m.qk_dwconv, m.qk(x_qk)).chunk(2, dim=1)
    v = _patches(_dw(m.v_dwconv, m.v(x)), ws)[:, None, None]          # b 1 1 N ws*ws*c
    if t0:
        q, k = _patches(q, ws), _patches(k, ws)
    else:
        k = m.k2_dwconv(m.k2(k)).flatten(2).transpose(1, 2)      # b N 2c
        q = m.q2_dwconv(m.q2(q)).flatten(2).transpose(1, 2)
    q, k = _unit_rows(q)[:, None, None], _unit_rows(k)[:, None, None]
    if k_hist is not None and v_hist is not None:
        k, v = torch.cat([k_hist, k], dim=1), torch.cat([v_hist, v], dim=1)
    nf = k.shape[1]
    if t0:
        o = v                                                    # T0:521-523: the aggregation result is discarded
    else:
        s = (q @ k.transpose(-1, -2)) * m.temperature            # b F 1 N N
        top = torch.topk(s, k=5, dim=-1).indices
        z = s * torch.zeros_like(s).scatter_(-1, top, 1.0) + s * _local_mask(h // ws, w // ws, s.device, s.dtype)
        o = _clipped_softmax(z) @ v
    o = _unpatch(o[:, :, 0], ws, c, h, w)
    o = m.project_out(o.reshape(b * nf, c, h, w)).reshape(b, nf, c, h, w)
    return o, k[:, -keep:], v[:, -keep:]


def _causal_history(m, x, variant, k_hist, v_hist):             # T1:627-662
    b, c, h, w = x.shape
    hd = m.num_heads
    xs, k_new, v_new = _state_align(m.spatial_aligner, x, variant, k_hist, v_hist)
    nf = xs.shape[1]
    k, v = _dw(m.kv_dwconv, m.kv(xs.reshape(b * nf, c, h, w))).chunk(2, dim=1)

    def rows(t):                                                 # '(b f) (head c) h w -> b head (f c) (h w)'
        return t.reshape(b, nf, hd, c // hd, h * w).permute(0, 2, 1, 3, 4).reshape(b, hd, nf * (c // hd), h * w)
    out, _, _ = _channel_attn(m.ChanAttn, x, _unit_rows(rows(k)), rows(v), keep_frames=1)
    return out, k_new, v_new


def _block(blk, x, variant, k_hist=None, v_hist=None):          # T1:804-811
    kc = vc = None
    t = blk.attention_type
    if t != "NoAttn":
        y = _layernorm(blk.norm1, x, conv_next=t in ("Channel", "ReducedAttn", "FHR"))
        if t == "Channel":
            o, _, _ = _channel_attn(blk.attn, y)
        elif t == "ReducedAttn":
            o = _reduced_attn(blk.attn, y)
        elif t == "FHR":
            o, kc, vc = _channel_attn(blk.attn, y, k_hist, v_hist, keep_frames=blk.attn.num_frames_tocache)
        else:
            o, kc, vc = _causal_history(blk.attn, y, variant, k_hist, v_hist)
        x = x + o
    y = _layernorm(blk.norm2, x, conv_next=True)
    return x + (_gated_ffw(blk.ffn, y) if blk.FFW_type == "GFFW" else _plain_ffw(blk.ffn, y)), kc, vc


def _level(lv, x, variant, k_hist=None, v_hist=None):           # T1:856-865
    kc = vc = None
    n = len(lv.transformer_blocks)
    for i, blk in enumerate(lv.transformer_blocks):
        last = i == n - 1
        x, kc, vc = _block(blk, x, variant, k_hist if last else None, v_hist if last else None)
    return x, kc, vc


def autograd_forward(net, pair: Tensor, k_cached: Optional[Sequence] = None, v_cached: Optional[Sequence] = None):
    """Differentiable ``Turtle*.forward`` (T1:1045-1132, T0:968-1050, TS:1049-1139) over ``net``'s parameters."""
    B, _, C, H, W = pair.shape
    variant = net.variant
    if k_cached is None:
        k_cached, v_cached = [None] * 8, [None] * 8
    img = torch.cat([pair[:, 0], pair[:, 1]], dim=1) if net.use_both_input else pair[:, 1]
    if variant == "super":
        img = F.interpolate(img, scale_factor=4, mode="bilinear")
        H, W = 4 * H, 4 * W
    img = net.check_image_size(img)
    current = img[:, C:] if net.use_both_input else img
    ks: List[Optional[Tensor]] = []
    vs: List[Optional[Tensor]] = []

    x = net.input_projection(img)
    e1, kc, vc = _level(net.encoder_level1, x, variant, k_cached[0], v_cached[0]); ks.append(kc); vs.append(vc)
    x = F.pixel_unshuffle(net.down1_2.body[0](e1), 2)
    e2, kc, vc = _level(net.encoder_level2, x, variant, k_cached[1], v_cached[1]); ks.append(kc); vs.append(vc)
    x = F.pixel_unshuffle(net.down2_3.body[0](e2), 2)
    e3, kc, vc = _level(net.encoder_level3, x, variant, k_cached[2], v_cached[2]); ks.append(kc); vs.append(vc)
    x = F.pixel_unshuffle(net.down3_4.body[0](e3), 2)

    lat = net.latent.transformer_blocks                          # T1:919-928
    for i, blk in enumerate(lat):
        if i == 0:
            x, k4, v4 = _block(blk, x, variant, k_cached[3], v_cached[3])
        elif i == len(lat) - 1:
            x, k5, v5 = _block(blk, x, variant, k_cached[4], v_cached[4])
        else:
            x, _, _ = _block(blk, x, variant)
    ks += [k4, k5]; vs += [v4, v5]

    x = F.pixel_shuffle(net.up4_3.body[0](x), 2)
    x = net.reduce_chan_level3(torch.cat([x, e3], 1))
    x, kc, vc = _level(net.decoder_level3, x, variant, k_cached[5], v_cached[5]); ks.append(kc); vs.append(vc)
    x = F.pixel_shuffle(net.up3_2.body[0](x), 2)
    x = net.reduce_chan_level2(torch.cat([x, e2], 1))
    x, kc, vc = _level(net.decoder_level2, x, variant, k_cached[6], v_cached[6]); ks.append(kc); vs.append(vc)
    x = F.pixel_shuffle(net.up2_1.body[0](x), 2)
    x = net.reduce_chan_level1(torch.cat([x, e1], 1))
    x, kc, vc = _level(net.decoder_level1, x, variant, k_cached[7], v_cached[7]); ks.append(kc); vs.append(vc)
    x, _, _ = _level(net.refinement, x, variant)
    out = net.ending(x) + current
    return out[:, :, :H, :W], ks, vs


# ----------------------------------------------------------------------------------------
# flat parameter / gradient storage and the bucketed all-reduce
# ----------------------------------------------------------------------------------------
class FlatParams:
    """All trainable parameters of ``net`` re-homed as views of one flat fp32 buffer (each tensor starts on a
    256-byte boundary), and a gradient buffer of the same layout whose views are installed as ``.grad``."""

    ALIGN = 64      # elements

    def __init__(self, net: torch.nn.Module):
        self.params = [p for p in net.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatParams needs fp32 parameters on one device")
            self.offsets.append(off)
            off += -(-p.numel() // self.ALIGN) * self.ALIGN
        self.numel = off
        self.data = torch.zeros(off, device=dev)
        self.grad = torch.zeros(off, device=dev)
        for p, o in zip(self.params, self.offsets):
            self.data[o:o + p.numel()].copy_(p.detach().reshape(-1))
            p.data = self.data[o:o + p.numel()].view(p.shape)
            p.grad = self.grad[o:o + p.numel()].view(p.shape)

    def zero_grad(self):
        self.grad.zero_()
        for p, o in zip(self.params, self.offsets):              # re-install views a user may have dropped
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + p.numel()].view(p.shape)


class GradBuckets:
    """Gradient mean over the data-parallel group: contiguous slices of the flat gradient buffer, cut in
    reverse construction order (the order backward produces them), each all-reduced asynchronously as soon as every
    gradient inside it has been accumulated.  ``finish()`` launches what is left (parameters the loss never touched,
    e.g. the unused k2/q2 convs of T0) and waits.  The sum -> mean division is left to the optimizer kernel
    (``scale``), so the collective is a plain SUM on both NCCL and gloo."""

    def __init__(self, flat: FlatParams, group=None, bucket_bytes: int = 32 << 20):
        import torch.distributed as dist
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.scale = 1.0 / self.world
        self.flat = flat
        cap = max(1, bucket_bytes // 4)
        # buckets are [lo, hi) element ranges of the flat buffer; parameter i belongs to exactly one bucket
        self.bucket_of = [0] * len(flat.params)
        self.ranges: List[List[int]] = []
        hi = flat.numel
        cur_lo = hi
        members = 0
        for i in reversed(range(len(flat.params))):
            lo = flat.offsets[i]
            if members and hi - lo > cap:
                self.ranges.append([cur_lo, hi])
                hi, members = cur_lo, 0
            cur_lo = lo
            self.bucket_of[i] = len(self.ranges)
            members += 1
        self.ranges.append([cur_lo, hi])
        self.members = [0] * len(self.ranges)
        for b in self.bucket_of:
            self.members[b] += 1
        self.pending = list(self.members)
        self.launched = [False] * len(self.ranges)
        self.works = []
        self.bytes_reduced = 0
        self.defer = False           # True: the hooks stay silent and finish() launches every bucket (graph replay)
        if self.world > 1:
            for i, p in enumerate(flat.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i):
        def hook(_p):
            if self.defer:
                return
            b = self.bucket_of[i]
            self.pending[b] -= 1
            if self.pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        if self.launched[b] or self.world == 1:
            return
        lo, hi = self.ranges[b]
        self.launched[b] = True
        self.bytes_reduced += 4 * (hi - lo)
        self.works.append(self.dist.all_reduce(self.flat.grad[lo:hi], group=self.group, async_op=True))

    def finish(self):
        for b in range(len(self.ranges)):
            self._launch(b)
        for w in self.works:
            w.wait()
        self.reset()

    def reset(self):
        self.works.clear()
        self.pending = list(self.members)
        self.launched = [False] * len(self.ranges)

    def reduce_loss(self, loss: Tensor) -> Tensor:
        """``reduce_loss_dict`` (BM:340-365): one-scalar ``reduce`` to rank 0, divided by the world size there."""
        if self.world == 1:
            return loss
        t = loss.detach().clone()
        self.dist.reduce(t, dst=0, group=self.group)
        if self.dist.get_rank(self.group) == 0:
            t /= self.world
        return t


# ----------------------------------------------------------------------------------------
# optimizer + loss scaler on the flat buffers
# ----------------------------------------------------------------------------------------
class FlatAdamW:
    """``torch.optim.AdamW`` (VRM:68-69) as one ``turtle_adamw_flat`` launch over the flat buffers."""

    def __init__(self, flat: FlatParams, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        from . import capi
        if flat.data.device.type != "cuda":
            raise RuntimeError("FlatAdamW runs on the sm_100a kernels of libturtle_b200.so only (no CPU optimizer path)")
        capi.load()
        self.capi = capi
        self.flat, self.lr, self.betas, self.eps, self.weight_decay = flat, lr, tuple(betas), eps, weight_decay
        self.exp_avg = torch.zeros_like(flat.data)
        self.exp_avg_sq = torch.zeros_like(flat.data)
        self.found_inf = torch.zeros(1, device=flat.data.device)
        self.steps = 0

    def step(self, grad_scale: float = 1.0, check_finite: bool = False) -> None:
        f = self.flat
        stream = torch.cuda.current_stream(f.data.device).cuda_stream
        found = None
        if check_finite:
            self.found_inf.zero_()
            self.capi.call("turtle_grad_check_finite", f.grad.data_ptr(), f.numel, self.found_inf.data_ptr(), stream)
            found = self.found_inf.data_ptr()
        self.capi.call("turtle_adamw_flat", f.data.data_ptr(), f.grad.data_ptr(), self.exp_avg.data_ptr(),
                       self.exp_avg_sq.data_ptr(), f.numel, self.lr, self.betas[0], self.betas[1], self.eps,
                       self.weight_decay, self.steps + 1, grad_scale, found, stream)
        if not check_finite:
            self.steps += 1

    def commit(self, skipped: bool) -> None:
        """After a checked step: the step count advances only if the update was applied (GradScaler.step)."""
        if not skipped:
            self.steps += 1


class LossScaler:
    """torch.cuda.amp.GradScaler's schedule (VRM:100-105): x0.5 on overflow, x2 after 2000 clean steps."""

    def __init__(self, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.scale, self.growth, self.backoff, self.interval = init_scale, growth_factor, backoff_factor, growth_interval
        self.clean = 0

    def update(self, found_inf: bool) -> None:
        if found_inf:
            self.scale *= self.backoff
            self.clean = 0
        else:
            self.clean += 1
            if self.clean == self.interval:
                self.scale *= self.growth
                self.clean = 0


class TrainStep:
    """``optimize_parameters`` (VRM:78-108): per-clip frame loop with BPTT through the caches, mean L1 over frames,
    backward, gradient mean over ranks, AdamW.  ``amp``: None (fp32), "fp16" (the reference's autocast + GradScaler)
    or "bf16" (autocast, no scaler)."""

    def __init__(self, net, optim: Optional[dict] = None, amp: Optional[str] = None, group=None,
                 bucket_bytes: int = 32 << 20, cuda_graph: bool = False):
        if amp not in (None, "fp16", "bf16"):
            raise ValueError(amp)
        optim = dict(optim or {})
        optim.pop("type", None)                                  # VRM:67: the yml's `type` is dropped, AdamW always
        self.net = net.train()
        self.flat = FlatParams(net)
        self.buckets = GradBuckets(self.flat, group, bucket_bytes)
        self.opt = FlatAdamW(self.flat, **optim)
        self.amp = amp
        self.scaler = LossScaler() if amp == "fp16" else None
        self.skipped_steps = 0
        # cuda_graph: after GRAPH_WARMUP eager steps the forward + backward of a step (several thousand launches issued
        # from Python; the step is bound by the host's launch rate) is captured once for the batch shape and replayed;
        # the gradient all-reduce and the optimizer stay outside the graph (the loss scale enters as a device scalar).
        self.cuda_graph = bool(cuda_graph)
        # overlap_in_graph: capture the bucketed all-reduces INSIDE the step graph (the post-accumulate hooks fire during
        # the captured backward, NCCL enqueues on its own stream forked from the capturing one, finish() joins it back),
        # so replayed steps overlap the gradient exchange with the backward pass exactly like eager steps do.  If the
        # capture of the collectives fails on this software stack, the step falls back to all-reducing after the graph.
        self.overlap_in_graph = True
        self._graph_has_allreduce = False
        self._graph = None
        self._eager_steps = 0
        self._scale_t = torch.ones((), device=self.flat.data.device)

    def loss_of_clip(self, lq: Tensor, gt: Tensor) -> Tensor:
        dev = lq.device.type
        ctx = torch.autocast(dev, dtype=torch.float16 if self.amp == "fp16" else torch.bfloat16) if self.amp \
            else torch.autocast(dev, enabled=False)
        n = lq.shape[1]
        total = 0
        k = v = None
        with ctx:
            for j in range(n):                                   # VRM:86-95
                pre = lq[:, j if j == 0 else j - 1]
                out, k, v = autograd_forward(self.net, torch.stack([pre, lq[:, j]], dim=1), k, v)
                total = total + F.l1_loss(out.float(), gt[:, j])
        return total / n

    GRAPH_WARMUP = 3

    def _replay(self, lq: Tensor, gt: Tensor) -> Tensor:
        if self._graph is None or self._g_lq.shape != lq.shape or self._g_lq.dtype != lq.dtype:
            self._g_lq, self._g_gt = lq.clone(), gt.clone()
            from . import capi
            n0 = capi.launch_count

            def capture(with_allreduce: bool):
                self.buckets.defer = not with_allreduce
                self.buckets.reset()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.flat.grad.zero_()
                    loss = self.loss_of_clip(self._g_lq, self._g_gt)
                    (loss * self._scale_t).backward()
                    if with_allreduce:
                        self.buckets.finish()                    # launches the stragglers and joins NCCL's stream
                    self._g_loss = loss.detach()
                return g

            want = self.overlap_in_graph and self.buckets.world > 1
            try:
                g = capture(want)
                self._graph_has_allreduce = want
            except Exception as e:                               # collectives not capturable here: exchange after the graph
                if not want:
                    raise
                import warnings
                warnings.warn(f"all-reduce could not be captured in the step graph ({type(e).__name__}: {e}); "
                              "falling back to the exchange after the graph")
                torch.cuda.synchronize()
                g = capture(False)
                self._graph_has_allreduce = False
            self.buckets.defer = True                            # replays: the hooks do not run, nothing to launch
            self._graph = g
            self._g_launches = capi.launch_count - n0          # our LayerNorm / depthwise launches inside the graph
            capi.launch_count = n0
        from . import capi
        self._g_lq.copy_(lq)
        self._g_gt.copy_(gt)
        self._graph.replay()
        capi.launch_count += self._g_launches
        return self._g_loss.clone()

    def step(self, lq: Tensor, gt: Tensor) -> Tensor:
        scale = self.scaler.scale if self.scaler else 1.0
        if self.cuda_graph and lq.is_cuda and self._eager_steps >= self.GRAPH_WARMUP:
            self._scale_t.fill_(scale)
            loss = self._replay(lq, gt)
        else:
            self._eager_steps += 1
            self.buckets.defer = False
            self.flat.zero_grad()
            loss = self.loss_of_clip(lq, gt)
            (loss * scale).backward()                            # bucket all-reduces start from the grad hooks
            loss = loss.detach()
        if not (self._graph is not None and self._graph_has_allreduce and self.buckets.defer):
            self.buckets.finish()                                # (replayed graphs that hold the collectives already did)
        self.opt.step(grad_scale=self.buckets.scale / scale, check_finite=self.scaler is not None)
        if self.scaler is not None:
            found = bool(self.opt.found_inf.item())
            self.opt.commit(found)
            self.scaler.update(found)
            self.skipped_steps += int(found)
        self.net.invalidate_packed_weights()                     # the inference engine repacks lazily
        return loss
