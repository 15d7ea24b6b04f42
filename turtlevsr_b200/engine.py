"""FrameEngine: the per-frame kernel schedule of the Turtle hot path.

One ``forward`` = one frame (T1:1045-1132).  Everything between the caller's NCHW frame and the
NCHW output runs in the hand-written kernels of libturtle_b200.so through the C ABI; torch is used
for device memory (workspace, parameters, history rings) and the stream only.  Activations are
fp32 channels-last.  There is no fallback path: CPU tensors or a missing library raise.

Schedule per block (each arrow is one kernel):

  ReducedAttn  LN -> 1x1(+b) -> dw3x3(+b)+GELU -> 1x1(+b)*beta+x
  FFW          LN -> 1x1(+b)+GELU -> 1x1(+b)*gamma+x
  GFFW         LN -> 1x1 -> dw3x3+gate -> 1x1+x
  Channel/FHR  LN -> 1x1 -> dw3x3 -> gram (per key segment) -> softmax -> fold W_o -> apply GEMM + x
               (FHR also pushes k-hat, v of the frame into its ring)
  CHM          LN -> SAB[ 1x1,dw (qk) ; 1x1,dw->ring (v) ; 1x1,window-reduce->ring (k) ; same (q) ;
                          select ; aggregate ; 1x1 ] -> 1x1,dw (kv over F frames) -> router as FHR
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import torch

from . import capi, ops
from .capi import GemmArgs
from .history import FhrRing, SabRing, resolve_ring

_NULL = None


def _ptr(t: Optional[torch.Tensor], off: int = 0):
    """Device address of element ``off`` of ``t`` that remembers its tensor (ops.DevPtr): the launches go through the
    torch custom-op layer (ops.py), which passes (tensor, byte offset) pairs to the dispatcher."""
    return ops.devptr(t, off)


class _Workspace:
    """Named, grow-only device scratch; steady-state frames allocate nothing."""

    def __init__(self, device, on_realloc=None):
        self.device = device
        self.bufs: Dict[str, torch.Tensor] = {}
        self.on_realloc = on_realloc          # called when a buffer that launches may already reference is replaced

    def get(self, name: str, *shape, dtype=torch.float32) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        b = self.bufs.get(name)
        if b is None or b.numel() < n or b.dtype != dtype:
            if b is not None and self.on_realloc is not None:
                self.on_realloc()
            b = torch.empty(max(n, 1), device=self.device, dtype=dtype)
            self.bufs[name] = b
        return b[:n].view(*shape)


class FrameEngine:
    def __init__(self, model):
        self.model = model
        self.packed: Dict[str, torch.Tensor] = {}
        self.ws: Optional[_Workspace] = None
        self.last_trace: Optional[dict] = None      # filled when model.record_trace is truthy
        # host-logic tests (CPU, no kernels): record the launch list instead of executing it
        self.dry_run = bool(getattr(model, "_dry_run", False))
        self.launch_log: List[str] = []
        if not self.dry_run:
            capi.load()                              # fail loudly now if the CUDA library is unavailable

        # per-kernel device timing (bench roofline leg): name -> [ms, launches, alg. bytes, flops]
        self.profile: Optional[dict] = None
        self.profile_shapes = False
        self._meta = (0, 0)
        # CUDA-graph replay (see _graph_forward): joint ring state -> captured frame
        self.graphs: Dict[tuple, tuple] = {}
        self._graph_seen = set()
        self.max_graphs = int(getattr(model, "cuda_graph_limit", 0) or self.MAX_GRAPHS)
        self.graph_captures = 0
        self.graph_replays = 0

    def _call(self, name, *args):
        if self.dry_run:
            self.launch_log.append(name)
            return
        if self.profile is None:
            ops.launch(name, *args)
            return
        # per-launch timing pass: straight through ctypes -- the dispatcher's extra host time per launch would sit between
        # the two events whenever the GPU drains the queue faster than the host fills it (kernels of a few tens of us)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.call(name, *args)
        e1.record()
        if self.profile_shapes and name in ("turtle_dwconv3x3", "turtle_layernorm"):
            if name == "turtle_dwconv3x3":      # (x, ldx, w, b, y, ldy, NB, H, W, C, fuse, layout, ws, rnd, stream)
                name += f"|C{args[9]}@{args[6]}x{args[7]}x{args[8]}|fuse{args[10]}|lay{args[11]}|rnd{args[13]}"
            else:                               # (x, ldx, w, b, y, ldy, P, C, rnd, stream)
                name += f"|C{args[7]}@{args[6]}|rnd{args[8]}"
        self.profile.setdefault("_events", []).append((name, e0, e1, self._meta))
        self._meta = (0, 0)

    def profile_begin(self, shapes: bool = False):
        self.profile = {}
        self.profile_shapes = shapes
        self._meta = (0, 0)          # a byte count noted while not profiling must not be billed to the first launch

    def profile_end(self) -> dict:
        """-> {kernel: dict(ms, launches, bytes, flops)} summed over the profiled frames."""
        torch.cuda.synchronize()
        out: Dict[str, dict] = {}
        for name, e0, e1, (by, fl) in self.profile.get("_events", []):
            d = out.setdefault(name, dict(ms=0.0, launches=0, bytes=0, flops=0))
            d["ms"] += e0.elapsed_time(e1)
            d["launches"] += 1
            d["bytes"] += by
            d["flops"] += fl
        self.profile = None
        return out

    def invalidate(self):
        self.packed.clear()
        self.drop_graphs()           # captured graphs hold the packed-weight addresses

    def drop_graphs(self):
        """Forget every captured frame: called when something a graph has baked in goes away -- packed weights,
        or a workspace buffer replaced by a larger one (a replayed graph would read and write freed memory)."""
        self.graphs.clear()
        self._graph_seen.clear()

    def _weights_fingerprint(self) -> int:
        """Changes whenever a parameter is updated in place (optimizer step, ``p.data = ...``, ``copy_``): the packed
        copies (fp16 / tap-major / TF32-rounded) and the graphs that hold their addresses are then stale.  A sum of
        autograd version counters and storage addresses -- a few hundred attribute reads per frame."""
        s = 0
        for p in self._plist:
            s += p._version + p.data_ptr()
        return s

    # ------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------
    def _param(self, name: str) -> Optional[torch.Tensor]:
        t = self._sd.get(name)
        return t

    def _w(self, name: str, kind: str = "raw") -> Optional[torch.Tensor]:
        """Device weight in the layout the kernels want (packed once, cached)."""
        if kind == "gemm16":          # fp16 copy of a 1x1 weight for the kind::f16 tensor-core path
            key = f"gemm16:{name}"
            t = self.packed.get(key)
            if t is None:
                t = self._sd[name].detach().float().reshape(self._sd[name].shape[0], -1).half().contiguous()
                self.packed[key] = t
            return t
        tf32 = self.mode == capi.TF32 and kind in ("gemm", "conv3")
        key = f"{kind}{'@tf32' if tf32 else ''}:{name}"
        t = self.packed.get(key)
        if t is not None:
            return t
        p = self._sd.get(name)
        if p is None:
            return None
        p = p.detach()
        if p.dtype != torch.float32:
            p = p.float()
        if kind == "raw" or kind == "gemm":   # vectors, beta/gamma, temperature / 1x1 conv [Cout,Cin,1,1]
            t = p.contiguous()
        elif kind == "dw":                # [C,1,k,k] -> tap-major [k*k, C]
            t = p.reshape(p.shape[0], -1).t().contiguous()
        elif kind == "dw16":              # the same, fp16 (taps of the fp16 depthwise kernel)
            t = p.reshape(p.shape[0], -1).t().contiguous().half()
        elif kind == "dw_lo" or kind == "dw_hi":     # halves of a depthwise weight (T0 q/k patches)
            h = p.shape[0] // 2
            q = p[:h] if kind == "dw_lo" else p[h:]
            t = q.reshape(h, -1).t().contiguous()
        elif kind == "dwgffw":            # GFFW depthwise taps per chunk of 32 gated channels: [hid/32][u1|u2][9][32] fp16
            hid = p.shape[0] // 2
            t = p.reshape(2, hid // 32, 32, 9).permute(1, 0, 3, 2).contiguous().half()
        elif kind == "conv3":             # [Cout,Cin,3,3] -> [Cout, 9*Cin] tap-major
            t = p.permute(0, 2, 3, 1).reshape(p.shape[0], -1).contiguous()
        elif kind == "conv316":           # the same, fp16 (kind::f16 im2col GEMM)
            t = p.permute(0, 2, 3, 1).reshape(p.shape[0], -1).contiguous().half()
        else:
            raise ValueError(kind)
        if tf32:
            # round the tensor-core weights to nearest TF32 once (the MMA would otherwise truncate them)
            t = ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
        self.packed[key] = t
        return t

    # ------------------------------------------------------------------------------------
    # kernel wrappers
    # ------------------------------------------------------------------------------------
    def gemm(self, segs, segw, Wt, out, ldo, P, Cout, bias=None, scale=None, act=0, res=None, ldres=0,
             im2col=0, geom=None, store=0, round_out=False, a16=False, o16=False, ln=None, wb=None):
        """segs: list of (ptr:int, lda:int).  round_out: the result feeds another tensor-core op.
        ln = (out_ptr, ld, norm_prefix): also emit fp16 LayerNorm(out rows) for the norm that reads them next.
        wb = (batches, stride between the weight matrices in elements, rows per batch): per-batch weights."""
        a = GemmArgs()
        a.mode = self.mode
        a.im2col = im2col
        a.P = P
        if geom is not None:
            a.B, a.H, a.W = geom
        a.Cout = Cout
        a.nseg = len(segs)
        a.segw = segw
        for i, (p, ld) in enumerate(segs):
            a.A[i] = p
            a.lda[i] = ld
        wt = _ptr(Wt) if isinstance(Wt, torch.Tensor) else Wt
        a.Wt = wt
        a.bias = _ptr(bias)
        a.scale = _ptr(scale)
        # the owners of the addresses above, for the custom-op layer (a ctypes struct keeps plain integers)
        a._ptrs = dict(A=[p for p, _ in segs], Wt=wt, bias=_ptr(bias), scale=_ptr(scale), res=res, out=out)
        a.act = act
        a.res = res
        a.ldres = ldres
        a.out = out
        a.ldo = ldo
        a.store = store
        a.round_out = 1 if (round_out and self.mode == capi.TF32 and not o16) else 0
        a.a_dtype = 1 if a16 else 0
        a.out_dtype = 1 if o16 else 0
        if wb is not None:
            a.w_batches, a.w_bstride, a.rows_per_batch = wb
        if ln is not None:
            a.ln_out, a.ld_ln = ln[0], ln[1]
            lw, lb = _ptr(self._w(ln[2] + "body.weight")), _ptr(self._w(ln[2] + "body.bias"))
            a.ln_w, a.ln_b = lw, lb
            a._ptrs.update(ln_out=ln[0], ln_w=lw, ln_b=lb)
        if self.profile is not None:
            K = (9 if im2col else len(segs)) * segw
            ea, eo = (2 if a16 else 4), (2 if o16 else 4)
            self._meta = (ea * (P * K + Cout * K) + eo * P * Cout + (4 * P * Cout if res else 0)
                          + (2 * P * Cout if ln is not None else 0), 2 * P * K * Cout)
        name = "turtle_gemm"
        if self.profile is not None:
            if self.profile_shapes:
                name = "turtle_gemm[conv3x3]" if im2col else "turtle_gemm[1x1]"
                Kt = (9 if im2col else len(segs)) * segw
                name += (f"|{Kt}->{Cout}@{P}" + ("+res" if res else "") + (f"/{len(segs)}seg" if len(segs) > 1 else "")
                         + ("|a16" if a16 else "") + ("|o16" if o16 else "") + ("|ln" if ln is not None else ""))
        self._call_gemm(name, a)

    def _gemm_launch(self, a):
        if self.profile is not None:                  # timing pass: no dispatcher between the events (see _call)
            launch = lambda: capi.call("turtle_gemm", C.byref(a), self.stream)
        else:
            launch = lambda: ops.launch_gemm(a, a._ptrs, self.stream)
        try:
            launch()
        except capi.TurtleKernelError as e:
            # the fused LayerNorm epilogue exists on the tensor-core kernel only: a shape that kernel does not cover
            # (sub-32 head widths of reduced configs) runs without it and the norm stays a launch of its own
            if e.code != capi.ENOTSUP or not a.ln_out:
                raise
            a.ln_out = None
            self._fused = None
            launch()

    def _call_gemm(self, tag, a):
        if self.dry_run:
            self.launch_log.append("turtle_gemm")
            return
        if self.profile is None:
            self._gemm_launch(a)
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self._gemm_launch(a)
        e1.record()
        self.profile.setdefault("_events", []).append((tag, e0, e1, self._meta))
        self._meta = (0, 0)

    def conv1x1(self, x, ldx, Cin, wname, out, ldo, P, Cout, **kw):
        self.gemm([(x, ldx)], Cin, self._w(wname, "gemm16" if kw.get("a16") else "gemm"), out, ldo, P, Cout, **kw)

    def half_path(self, *chans) -> bool:
        """fp16 storage for the FFN-side intermediates (LN output, wide hidden, gated hidden): tf32 mode only and
        only for shapes the fp16 kernels cover (32-aligned channel counts, LN widths 64..512)."""
        return (self.mode == capi.TF32 and self.use_half and not self.dry_run
                and chans[0] in (64, 128, 256, 512) and all(ch % 32 == 0 for ch in chans))

    def ln_fusable(self, c: int, prefix: str) -> bool:
        """Can the GEMM that updates the residual stream also emit fp16 LayerNorm(x) for norm ``prefix``?
        (tensor-core mode, whole row in one n-group, WithBias norm)"""
        return (self.mode == capi.TF32 and self.use_half and self.fuse_ln and not self.dry_run and c in (64, 128, 256)
                and (prefix + "body.bias") in self._sd)

    def ln_target(self, P: int, c: int, prefix: str, b0: int = 0, Pimg: int = 0, avoid: Optional[torch.Tensor] = None):
        """-> (ptr, ld, prefix) of the fp16 buffer the fused LayerNorm writes (rows b0*Pimg.. of it).
        ``avoid``: a map the producing kernel is still reading while it writes (the fused GFFW kernel reads its fp16 input
        tile by tile, halo included): the output then goes to the second buffer."""
        y = self.ws.get("xn16", P, c, dtype=torch.float16)
        if avoid is not None and y.data_ptr() == avoid.data_ptr():
            y = self.ws.get("xn16b", P, c, dtype=torch.float16)
        self._fused = (prefix, y)
        return (_ptr(y, b0 * Pimg * c), c, prefix)

    def take_fused(self, prefix: str):
        f, self._fused = self._fused, None
        if f is not None and f[0] == prefix:
            return f[1]
        return None

    def layernorm(self, x: torch.Tensor, pre: str, C_: int, P: int, half: bool = False) -> torch.Tensor:
        if half:
            y = self.take_fused(pre)
            if y is not None:
                return y
            y = self.ws.get("xn16", P, C_, dtype=torch.float16)
            self._meta = (6 * P * C_, 0)
            self._call("turtle_layernorm", _ptr(x), C_, _ptr(self._w(pre + "body.weight")),
                       _ptr(self._w(pre + "body.bias")), _ptr(y), C_, P, C_, 2, self.stream)
            return y
        self._fused = None
        y = self.ws.get("xn", P, C_)
        self._meta = (8 * P * C_, 0)
        self._call("turtle_layernorm", _ptr(x), C_, _ptr(self._w(pre + "body.weight")), _ptr(self._w(pre + "body.bias")),
             _ptr(y), C_, P, C_, self.rnd, self.stream)
        return y

    def dwconv(self, x, ldx, wname, bname, y, ldy, NB, H, W, Cc, fuse=0, layout=0, ws=1, wkind="dw", rnd=None, boff=0):
        """bname: bias parameter name or None (absent parameters are skipped); boff: first channel of it to use."""
        rnd = (self.rnd if layout == 0 else 0) if rnd is None else rnd      # ring / patch rows stay unrounded
        self._meta = (4 * NB * H * W * (Cc + (Cc // 2 if fuse == 2 else Cc)), 2 * 9 * NB * H * W * Cc)
        self._call("turtle_dwconv3x3", x, ldx, _ptr(self._w(wname, wkind)), _ptr(self._w(bname), boff) if bname else None, y,
                   ldy, NB, H, W, Cc, fuse, layout, ws, rnd, self.stream)

    # ------------------------------------------------------------------------------------
    # feed-forwards (x updated in place)
    # ------------------------------------------------------------------------------------
    def gated_ffw(self, pre, xn, x, P, c, H, W, B, ln=None):
        hid2 = self._sd[pre + "project_in.weight"].shape[0]
        hid = hid2 // 2
        if (xn.dtype == torch.float16 and self.fuse_gffw and c in (64, 128, 256) and hid % 32 == 0
                and (pre + "project_in.bias") not in self._sd and (pre + "dwconv.bias") not in self._sd
                and (pre + "project_out.bias") not in self._sd):
            # one kernel: the 5c-wide hidden map never leaves the SM (csrc/gffw_fused.cu)
            lt = self.ln_target(P, c, ln, avoid=xn) if ln else None
            self._meta = (P * c * (2 + 4 + 4 + (2 if ln else 0)) + 2 * 3 * hid * c, 2 * P * c * 3 * hid + 2 * 9 * P * hid2)
            self._call("turtle_gffw_fused", _ptr(xn), _ptr(self._w(pre + "project_in.weight", "gemm16")),
                       _ptr(self._w(pre + "dwconv.weight", "dwgffw")), _ptr(self._w(pre + "project_out.weight", "gemm16")),
                       _ptr(x), lt[0] if lt else None, _ptr(self._w(ln + "body.weight")) if lt else None,
                       _ptr(self._w(ln + "body.bias")) if lt else None, B, H, W, c, hid, self.stream)
            return
        if xn.dtype == torch.float16:       # fp16 intermediates, kind::f16 MMAs
            t = self.ws.get("wide16", P, hid2, dtype=torch.float16)
            self.conv1x1(_ptr(xn), c, c, pre + "project_in.weight", _ptr(t), hid2, P, hid2,
                         bias=self._w(pre + "project_in.bias"), a16=True, o16=True)
            if (self.gffw_tail and c in (64, 128, 256) and hid % 32 == 0 and hid * 2 <= 5 * c
                    and (pre + "dwconv.bias") not in self._sd and (pre + "project_out.bias") not in self._sd):
                # depthwise + gate as the A-producer of project_out: the gated map never reaches HBM (csrc/gffw_fused.cu)
                lt = self.ln_target(P, c, ln) if ln else None
                self._meta = (P * (2 * hid2 + c * (4 + 4 + (2 if ln else 0))) + 2 * hid * c, 2 * P * c * hid + 2 * 9 * P * hid2)
                self._call("turtle_gffw_tail", _ptr(t), _ptr(self._w(pre + "dwconv.weight", "dwgffw")),
                           _ptr(self._w(pre + "project_out.weight", "gemm16")), _ptr(x), lt[0] if lt else None,
                           _ptr(self._w(ln + "body.weight")) if lt else None, _ptr(self._w(ln + "body.bias")) if lt else None,
                           B, H, W, c, hid, self.stream)
                return
            g = self.ws.get("dw16", P, hid, dtype=torch.float16)
            self._meta = (2 * P * (hid2 + hid), 2 * 9 * P * hid2)
            self._call("turtle_dwconv3x3", _ptr(t), hid2, _ptr(self._w(pre + "dwconv.weight", "dw16")),
                       _ptr(self._w(pre + "dwconv.bias")), _ptr(g), hid, B, H, W, hid2, 2, 0, 1, 2, self.stream)
            self.conv1x1(_ptr(g), hid, hid, pre + "project_out.weight", _ptr(x), c, P, c,
                         bias=self._w(pre + "project_out.bias"), res=_ptr(x), ldres=c, a16=True,
                         ln=self.ln_target(P, c, ln) if ln else None)
            return
        t = self.ws.get("wide", P, hid2)
        self.conv1x1(_ptr(xn), c, c, pre + "project_in.weight", _ptr(t), hid2, P, hid2,
                     bias=self._w(pre + "project_in.bias"))
        g = self.ws.get("dw", P, hid)
        self.dwconv(_ptr(t), hid2, pre + "dwconv.weight", pre + "dwconv.bias", _ptr(g), hid, B, H, W, hid2, fuse=2)
        self.conv1x1(_ptr(g), hid, hid, pre + "project_out.weight", _ptr(x), c, P, c,
                     bias=self._w(pre + "project_out.bias"), res=_ptr(x), ldres=c)

    def plain_ffw(self, pre, xn, x, P, c, ln=None):
        if xn.dtype == torch.float16:
            t = self.ws.get("wide16", P, 2 * c, dtype=torch.float16)
            self.conv1x1(_ptr(xn), c, c, pre + "conv4.weight", _ptr(t), 2 * c, P, 2 * c,
                         bias=self._w(pre + "conv4.bias"), act=capi.ACT_GELU, a16=True, o16=True)
            self.conv1x1(_ptr(t), 2 * c, 2 * c, pre + "conv5.weight", _ptr(x), c, P, c, bias=self._w(pre + "conv5.bias"),
                         scale=self._w(pre + "gamma"), res=_ptr(x), ldres=c, a16=True,
                         ln=self.ln_target(P, c, ln) if ln else None)
            return
        t = self.ws.get("wide", P, 2 * c)
        self.conv1x1(_ptr(xn), c, c, pre + "conv4.weight", _ptr(t), 2 * c, P, 2 * c, bias=self._w(pre + "conv4.bias"),
                     act=capi.ACT_GELU, round_out=True)
        self.conv1x1(_ptr(t), 2 * c, 2 * c, pre + "conv5.weight", _ptr(x), c, P, c, bias=self._w(pre + "conv5.bias"),
                     scale=self._w(pre + "gamma"), res=_ptr(x), ldres=c)

    # ------------------------------------------------------------------------------------
    # attentions (x updated in place: x += attn(xn))
    # ------------------------------------------------------------------------------------
    def reduced_attn(self, pre, xn, x, P, c, H, W, B, ln=None):
        if xn.dtype == torch.float16:
            t = self.ws.get("wide16", P, 2 * c, dtype=torch.float16)
            self.conv1x1(_ptr(xn), c, c, pre + "conv1.weight", _ptr(t), 2 * c, P, 2 * c,
                         bias=self._w(pre + "conv1.bias"), a16=True, o16=True)
            u = self.ws.get("dw16", P, 2 * c, dtype=torch.float16)
            self._meta = (2 * P * 4 * c, 2 * 9 * P * 2 * c)
            self._call("turtle_dwconv3x3", _ptr(t), 2 * c, _ptr(self._w(pre + "conv2.weight", "dw16")),
                       _ptr(self._w(pre + "conv2.bias")), _ptr(u), 2 * c, B, H, W, 2 * c, 1, 0, 1, 2, self.stream)
            self.conv1x1(_ptr(u), 2 * c, 2 * c, pre + "conv3.weight", _ptr(x), c, P, c, bias=self._w(pre + "conv3.bias"),
                         scale=self._w(pre + "beta"), res=_ptr(x), ldres=c, a16=True,
                         ln=self.ln_target(P, c, ln) if ln else None)
            return
        t = self.ws.get("wide", P, 2 * c)
        self.conv1x1(_ptr(xn), c, c, pre + "conv1.weight", _ptr(t), 2 * c, P, 2 * c, bias=self._w(pre + "conv1.bias"))
        u = self.ws.get("dw", P, 2 * c)
        self.dwconv(_ptr(t), 2 * c, pre + "conv2.weight", pre + "conv2.bias", _ptr(u), 2 * c, B, H, W, 2 * c, fuse=1)
        self.conv1x1(_ptr(u), 2 * c, 2 * c, pre + "conv3.weight", _ptr(x), c, P, c, bias=self._w(pre + "conv3.bias"),
                     scale=self._w(pre + "beta"), res=_ptr(x), ldres=c)

    def channel_attn(self, pre, xn, x, B, H, W, c, heads, hist_segs=None, ring: Optional[FhrRing] = None,
                     ring_slot: int = -1, ln=None, force16: bool = False):
        """ChannelAttention / FrameHistoryRouter.

        hist_segs: per batch element a list of key/value history segments (oldest first), each
        ``dict(k=ptr, ldk=, khs=, v=ptr, ldv=, vhs=, prenorm=bool)``.  The frame's own q/k/v are
        appended as the last segment.  With ``ring`` the normalised key rows and value rows of the
        frame are pushed into slot ``ring_slot`` (T1:286).

        When ``xn`` is fp16 (tf32 mode, no ring) q/k/v and the folded weight stay fp16 end to end: the Gram and
        the apply GEMM run as kind::f16 MMAs (history segments must then be fp16 too)."""
        Pimg = H * W
        P = B * Pimg
        ch = c // heads
        a16 = xn.dtype == torch.float16
        h16 = a16 or force16            # force16: fp32 (TF32) input rows, everything downstream of the qkv conv in fp16
        dt = torch.float16 if h16 else torch.float32
        es = 2 if h16 else 4
        qkv = self.ws.get("wide16" if h16 else "wide", P, 3 * c, dtype=dt)
        self.conv1x1(_ptr(xn), c, c, pre + "qkv.weight", _ptr(qkv), 3 * c, P, 3 * c, bias=self._w(pre + "qkv.bias"),
                     a16=a16, o16=h16)
        qd = self.ws.get("dw16" if h16 else "dw", P, 3 * c, dtype=dt)
        if h16:
            self._meta = (2 * P * 6 * c, 2 * 9 * P * 3 * c)
            self._call("turtle_dwconv3x3", _ptr(qkv), 3 * c, _ptr(self._w(pre + "qkv_dwconv.weight", "dw16")),
                       _ptr(self._w(pre + "qkv_dwconv.bias")), _ptr(qd), 3 * c, B, H, W, 3 * c, 0, 0, 1, 2, self.stream)
        else:
            self.dwconv(_ptr(qkv), 3 * c, pre + "qkv_dwconv.weight", pre + "qkv_dwconv.bias", _ptr(qd), 3 * c, B, H, W,
                        3 * c)
        # pixel splits of the Gram: the batch elements share the CTA budget (gram / softmax / fold run once for all of them)
        nsplit = max(1, min((Pimg + 255) // 256, max(1, self.gram_ctas // (heads * B))))
        temp = self._w(pre + "temperature")
        Wo = self._w(pre + "project_out.weight")
        gmode = 2 if h16 else self.mode
        # key / value segments per batch element: history (oldest first), then the frame's own k / v
        allsegs = []
        for b in range(B):
            segs = list(hist_segs[b]) if hist_segs is not None else []
            base = b * Pimg * 3 * c
            segs.append(dict(k=_ptr(qd, base + c), ldk=3 * c, khs=ch, v=_ptr(qd, base + 2 * c), ldv=3 * c, vhs=ch,
                             prenorm=False))
            allsegs.append(segs)
        S = len(allsegs[0])
        # one launch per segment covers every batch element when element b of a segment sits at a fixed stride from
        # element 0 (true for the rings, the CHM frame stack and the workspace maps); otherwise one launch per (b, segment)
        kbs = []
        for s_ in range(S):
            d = (int(allsegs[1][s_]["k"]) - int(allsegs[0][s_]["k"])) if B > 1 else 0
            ok = d >= 0 and d % es == 0 and all(int(allsegs[b][s_]["k"]) - int(allsegs[0][s_]["k"]) == b * d and
                                                 allsegs[b][s_]["ldk"] == allsegs[0][s_]["ldk"] and
                                                 allsegs[b][s_]["khs"] == allsegs[0][s_]["khs"] for b in range(B))
            kbs.append(d // es if ok else None)
        gpart = self.ws.get("gram", B, S, nsplit, heads, ch, ch)
        sqq = self.ws.get("sqq", B, S, nsplit, c)
        sqk = self.ws.get("sqk", B, S, nsplit, c)
        g_bs, s_bs = S * nsplit * heads * ch * ch, S * nsplit * c
        for s_ in range(S):
            sg = allsegs[0][s_]
            if kbs[s_] is not None:
                self._meta = (2 * es * P * c, 2 * P * c * ch)
                self._call("turtle_chan_gram_b", _ptr(qd), 3 * c, ch, Pimg * 3 * c, sg["k"], sg["ldk"], sg["khs"], kbs[s_],
                           Pimg, heads, ch, nsplit, _ptr(gpart[0, s_]), _ptr(sqq[0, s_]), _ptr(sqk[0, s_]), g_bs, s_bs, B,
                           gmode, self.stream)
            else:
                for b in range(B):
                    sgb = allsegs[b][s_]
                    self._meta = (2 * es * Pimg * c, 2 * Pimg * c * ch)
                    self._call("turtle_chan_gram", _ptr(qd, b * Pimg * 3 * c), 3 * c, ch, sgb["k"], sgb["ldk"], sgb["khs"],
                               Pimg, heads, ch, nsplit, _ptr(gpart[b, s_]), _ptr(sqq[b, s_]), _ptr(sqk[b, s_]), gmode,
                               self.stream)
        flags = self._flags([1 if sg["prenorm"] else 0 for sg in allsegs[0]])
        Pm = self.ws.get("attnP", B, heads, ch, S * ch)
        inv = self.ws.get("invk", B, S, c)
        self._call("turtle_chan_softmax_b", _ptr(gpart), _ptr(sqq), _ptr(sqk), _ptr(flags), _ptr(temp), S, nsplit, heads, ch,
                   _ptr(Pm), _ptr(inv), g_bs, s_bs, B, self.stream)
        M = self.ws.get("attnM16" if h16 else "attnM", B, c, S * c, dtype=dt)
        self._call("turtle_chan_fold_b", _ptr(Pm), _ptr(Wo), S, heads, ch, _ptr(M), 2 if h16 else self.rnd, B, self.stream)
        # apply: x += (W_o . blockdiag(P_b)) v_b.  One launch with per-image weight matrices when every 128-row tile lies
        # inside one image and image b's value rows follow image b-1's in memory (workspace maps and the FHR ring do; the
        # CHM frame stack does not); otherwise one launch per image
        batched = (B > 1 and self.mode == capi.TF32 and Pimg % 128 == 0 and not self.dry_run and
                   all(int(allsegs[b][s_]["v"]) - int(allsegs[0][s_]["v"]) == b * Pimg * allsegs[0][s_]["ldv"] * es and
                       allsegs[b][s_]["ldv"] == allsegs[0][s_]["ldv"] and allsegs[b][s_]["vhs"] == allsegs[0][s_]["vhs"]
                       for b in range(B) for s_ in range(S)))
        if batched:
            vsegs = [(sg["v"] + es * h * sg["vhs"], sg["ldv"]) for sg in allsegs[0] for h in range(heads)]
            try:
                self.gemm(vsegs, ch, M, _ptr(x), c, P, c, bias=self._w(pre + "project_out.bias"), res=_ptr(x), ldres=c,
                          a16=h16, ln=self.ln_target(P, c, ln) if ln else None, wb=(B, c * S * c, Pimg))
            except capi.TurtleKernelError as e:
                if e.code != capi.ENOTSUP:
                    raise
                batched = False
        for b in range(B):
            segs = allsegs[b]
            if batched:
                break
            vsegs = [(sg["v"] + es * h * sg["vhs"], sg["ldv"]) for sg in segs for h in range(heads)]
            xb = _ptr(x, b * Pimg * c)
            self.gemm(vsegs, ch, M[b], xb, c, Pimg, c, bias=self._w(pre + "project_out.bias"), res=xb, ldres=c, a16=h16,
                      ln=self.ln_target(P, c, ln, b, Pimg) if ln else None)
        for b in range(B):
            if ring is not None:
                base = b * Pimg * 3 * c
                self._call("turtle_scale_cols", _ptr(qd, base + c), 3 * c, ch, _ptr(inv[b, S - 1]),
                           ring.slot_ptr(ring.kbuf, b, ring_slot), ring.ld, ring.head_stride, Pimg, heads, ch,
                           self.stream)
                self._call("turtle_scale_cols", _ptr(qd, base + 2 * c), 3 * c, ch, None,
                           ring.slot_ptr(ring.vbuf, b, ring_slot), ring.ld, ring.head_stride, Pimg, heads, ch,
                           self.stream)
        if self.trace is not None:
            self.trace.setdefault(pre, []).append(dict(qkv_dw=qd.clone()))

    def _flags(self, vals: List[int]) -> torch.Tensor:
        key = "flags:" + "".join(map(str, vals))
        t = self.packed.get(key)
        if t is None:
            t = torch.tensor(vals, dtype=torch.int32, device=self.device)
            self.packed[key] = t
        return t

    def fhr(self, pre, xn, x, B, H, W, c, heads, keep, k_in, v_in, ln=None):
        """FrameHistoryRouter with its ring (latent blocks 0 and -1, T1:243-286)."""
        Pimg, ch = H * W, c // heads
        ring = resolve_ring(k_in, v_in)
        if ring is None:
            if k_in is not None and v_in is not None:
                ring = FhrRing.adopt(k_in, v_in, keep, ch, self.device)
            else:
                ring = FhrRing(B, Pimg, heads, ch, keep, self.device)
        if ring.geometry() != (B, Pimg, heads, ch):
            # the reference fails in torch.cat here (T1:272-273); without the check the kernels would run off the ring
            raise ValueError(f"{pre}: history caches were built for (batch, pixels, heads, head width) = "
                             f"{ring.geometry()}, this frame has {(B, Pimg, heads, ch)}")
        slot = ring.begin_push()
        hist = []
        for b in range(B):
            segs = []
            for t in range(ring.first_live, ring.pos + 1):
                segs.append(dict(k=ring.slot_ptr(ring.kbuf, b, t), ldk=ring.ld, khs=ring.head_stride,
                                 v=ring.slot_ptr(ring.vbuf, b, t), ldv=ring.ld, vhs=ring.head_stride, prenorm=True))
            hist.append(segs)
        self.channel_attn(pre, xn, x, B, H, W, c, heads, hist_segs=hist, ring=ring, ring_slot=slot, ln=ln)
        # the reference returns cat(history, new)[-K:]; the window after commit is exactly that
        ring.commit()
        return ring.views()

    def chm(self, pre, xn, x, B, H, W, c, heads, scale_patch, keep, k_in, v_in, ln=None):
        """CausalHistoryModel (T1:627-662) = StateAlignBlock (T1:548-610 / T0:459-533) + router."""
        t0 = self.model.variant == "t0"
        sa = pre + "spatial_aligner."
        ws_ = 2 * scale_patch
        Pimg = H * W
        P = B * Pimg
        Hg, Wg = H // ws_, W // ws_
        N = Hg * Wg
        Dk = ws_ * ws_ * c if t0 else 2 * c
        Dv = ws_ * ws_ * c
        ring = resolve_ring(k_in, v_in)
        if ring is None:
            if k_in is not None and v_in is not None:
                ring = SabRing.adopt(k_in, v_in, keep, self.device)
            else:
                ring = SabRing(B, N, Dk, Dv, keep, self.device)
        if ring.geometry() != (B, N, Dk, Dv):
            # the reference fails in torch.cat here (T1:581-582)
            raise ValueError(f"{pre}: history caches were built for (batch, patches, key width, value width) = "
                             f"{ring.geometry()}, this frame has {(B, N, Dk, Dv)}")
        slot = ring.begin_push()
        first = ring.first_live
        F_ = ring.count + 1

        # --- q/k/v feature maps -----------------------------------------------------------
        src = xn
        if t0:
            if c % 4 != 0:           # same failure as positionalencoding2d, T0:424-426
                raise ValueError("Cannot use sin/cos positional encoding with odd dimension (got dim={:d})".format(c))
            src = self.ws.get("xpe", P, c)
            self._call("turtle_add_posenc", _ptr(xn), _ptr(src), B, H, W, c, self.stream)
        # selection front end in fp16 (tensor-core mode): qk map, its depthwise, the k2 / q2 maps; sums stay fp32
        s16 = (not t0) and self.sab_front_half and self.half_path(c, 2 * c)
        sdt = torch.float16 if s16 else torch.float32
        qk = self.ws.get("wide16" if s16 else "wide", P, 2 * c, dtype=sdt)
        self.conv1x1(_ptr(src), c, c, sa + "qk.weight", _ptr(qk), 2 * c, P, 2 * c, bias=self._w(sa + "qk.bias"), o16=s16)
        qkd = self.ws.get("dw16" if s16 else "dw", P, 2 * c, dtype=sdt)
        vt = self.ws.get("sab_v", P, c)
        self.conv1x1(_ptr(xn), c, c, sa + "v.weight", _ptr(vt), c, P, c, bias=self._w(sa + "v.bias"))
        qn = self.ws.get("sab_qn", B, N, Dk)
        if not t0 and s16:
            self._meta = (2 * P * 4 * c, 2 * 9 * P * 2 * c)
            self._call("turtle_dwconv3x3", _ptr(qk), 2 * c, _ptr(self._w(sa + "qk_dwconv.weight", "dw16")),
                       _ptr(self._w(sa + "qk_dwconv.bias")), _ptr(qkd), 2 * c, B, H, W, 2 * c, 0, 0, 1, 2, self.stream)
            red = self.ws.get("wide16", P, 2 * c, dtype=torch.float16)            # qk no longer needed
            self.conv1x1(_ptr(qkd, c), 2 * c, c, sa + "k2.weight", _ptr(red), 2 * c, P, 2 * c, bias=self._w(sa + "k2.bias"),
                         a16=True, o16=True)
            self._call("turtle_sab_window_reduce_h16", _ptr(red), 2 * c, _ptr(self._w(sa + "k2_dwconv.weight", "dw")),
                       _ptr(self._w(sa + "k2_dwconv.bias")), _ptr(ring.kbuf[:, slot]), ring.kbuf.stride(0), B, H, W, 2 * c,
                       ws_, self.stream)
            self.conv1x1(_ptr(qkd), 2 * c, c, sa + "q2.weight", _ptr(red), 2 * c, P, 2 * c, bias=self._w(sa + "q2.bias"),
                         a16=True, o16=True)
            self._call("turtle_sab_window_reduce_h16", _ptr(red), 2 * c, _ptr(self._w(sa + "q2_dwconv.weight", "dw")),
                       _ptr(self._w(sa + "q2_dwconv.bias")), _ptr(qn), N * Dk, B, H, W, 2 * c, ws_, self.stream)
        elif not t0:
            self.dwconv(_ptr(qk), 2 * c, sa + "qk_dwconv.weight", sa + "qk_dwconv.bias", _ptr(qkd), 2 * c, B, H, W, 2 * c)
            red = self.ws.get("wide", P, 2 * c)            # qk no longer needed
            # k: 1x1 c->2c on the k half, then window reduce + normalise straight into the ring slot
            self.conv1x1(_ptr(qkd, c), 2 * c, c, sa + "k2.weight", _ptr(red), 2 * c, P, 2 * c, bias=self._w(sa + "k2.bias"))
            self._call("turtle_sab_window_reduce", _ptr(red), 2 * c, _ptr(self._w(sa + "k2_dwconv.weight", "dw")),
                       _ptr(self._w(sa + "k2_dwconv.bias")), _ptr(ring.kbuf[:, slot]), ring.kbuf.stride(0), B, H, W, 2 * c,
                       ws_, self.stream)
            self.conv1x1(_ptr(qkd), 2 * c, c, sa + "q2.weight", _ptr(red), 2 * c, P, 2 * c, bias=self._w(sa + "q2.bias"))
            self._call("turtle_sab_window_reduce", _ptr(red), 2 * c, _ptr(self._w(sa + "q2_dwconv.weight", "dw")),
                       _ptr(self._w(sa + "q2_dwconv.bias")), _ptr(qn), N * Dk, B, H, W, 2 * c, ws_, self.stream)
        # aggregation on the tensor cores (csrc/sab_agg_tc.cu) reads an fp16 copy of the value rows kept next to the ring
        agg_tc = (not t0 and self.mode == capi.TF32 and getattr(self.model, "sab_agg_tc", True) and c % 32 == 0
                  and Dv % 256 == 0)
        v16 = ring.shadow() if agg_tc else None
        v16_new = False                  # the new slot's copy was written by the depthwise kernel itself
        for b in range(B):
            # v: depthwise 3x3 written directly as dilated patch rows into the ring slot
            if agg_tc:
                try:
                    self._meta = (4 * Pimg * c * 2 + 2 * Pimg * c, 2 * 9 * Pimg * c)
                    self._call("turtle_dwconv3x3_patch_rows", _ptr(vt, b * Pimg * c), c, _ptr(self._w(sa + "v_dwconv.weight", "dw")),
                               _ptr(self._w(sa + "v_dwconv.bias")), _ptr(ring.vbuf[b, slot]), _ptr(v16[b, slot]), 1, H, W, c, ws_,
                               self.stream)
                    v16_new = True
                    continue
                except capi.TurtleKernelError as e:
                    if e.code != capi.ENOTSUP:
                        raise
                    v16_new = False
            self.dwconv(_ptr(vt, b * Pimg * c), c, sa + "v_dwconv.weight", sa + "v_dwconv.bias", _ptr(ring.vbuf[b, slot]), c,
                        1, H, W, c, layout=1, ws=ws_)
            if t0:
                self.dwconv(_ptr(qk, b * Pimg * 2 * c), 2 * c, sa + "qk_dwconv.weight", sa + "qk_dwconv.bias", _ptr(qn[b]),
                            c, 1, H, W, c, layout=1, ws=ws_, wkind="dw_lo")
                self.dwconv(_ptr(qk, b * Pimg * 2 * c + c), 2 * c, sa + "qk_dwconv.weight", sa + "qk_dwconv.bias",
                            _ptr(ring.kbuf[b, slot]), c, 1, H, W, c, layout=1, ws=ws_, wkind="dw_hi", boff=c)
                self._call("turtle_sab_patch_normalize", _ptr(qn[b]), N, Dk, self.stream)
                self._call("turtle_sab_patch_normalize", _ptr(ring.kbuf[b, slot]), N, Dk, self.stream)

        # --- selection + aggregation over the F live frames ------------------------------------
        # router side in fp16 (tensor-core mode): aligned frames, their kv maps and the inner channel attention
        r16 = (not t0) and self.half_path(c, 2 * c) and c // heads == 64
        dt16 = torch.float16 if r16 else torch.float32
        agg = self.ws.get("sab_agg16" if r16 else "sab_agg", B, F_, Pimg, c, dtype=dt16)
        idx = self.ws.get("sab_idx", B, F_, N, capi.SAB_SLOTS, dtype=torch.int32)
        wgt = self.ws.get("sab_wgt", B, F_, N, capi.SAB_SLOTS)
        temp = self._w(sa + "temperature")
        if agg_tc:
            lo = slot if ring.count == 0 else ring.first_live
            have = ring.v16_lo if ring.v16_lo is not None else slot
            # history slots without a copy yet (adopted caches), then the new frame unless the depthwise kernel wrote it
            for s_ in list(range(lo, min(have, slot))) + ([] if v16_new else [slot]):
                for b in range(B):
                    self._meta = (6 * N * Dv, 0)
                    self._call("turtle_cast_f16", _ptr(ring.vbuf[b, s_]), _ptr(v16[b, s_]), N * Dv, self.stream)
            ring.v16_lo = min(lo, have)
            agg_ws = self.ws.get("sab_aggws", (capi.load().turtle_sab_aggregate_tc_workspace(F_, Hg, Wg) + 3) // 4)
        else:
            ring.v16_lo = None
        for b in range(B):
            kf = _ptr(ring.kbuf[b, first])
            vf = _ptr(ring.vbuf[b, first])
            if not t0:
                self._meta = (4 * (N * Dk * (F_ + 1) + 2 * F_ * N * capi.SAB_SLOTS), 2 * F_ * N * N * Dk)
                if self.mode == capi.TF32 and Dk % 32 == 0 and N >= 5 and not self.dry_run:
                    wsb = capi.load().turtle_sab_select_tc_workspace(F_, N, Dk)
                    wsp = self.ws.get("sab_tcws", (wsb + 3) // 4)
                    self._call("turtle_sab_select_tc", _ptr(qn[b]), kf, N * Dk, F_, Hg, Wg, Dk, _ptr(temp), 0,
                               _ptr(idx[b]), _ptr(wgt[b]), _ptr(wsp), self.stream)
                else:
                    self._call("turtle_sab_select", _ptr(qn[b]), kf, N * Dk, F_, Hg, Wg, Dk, _ptr(temp), 0,
                               _ptr(idx[b]), _ptr(wgt[b]), self.mode, self.stream)
            if agg_tc:
                self._meta = (F_ * N * Dv * (2 + (2 if r16 else 4)) + 8 * F_ * N * capi.SAB_SLOTS, 2 * F_ * N * 46 * Dv)
                self._call("turtle_sab_aggregate_tc", _ptr(idx[b]), _ptr(wgt[b]), _ptr(v16[b, first]), 1, N * Dv, _ptr(agg[b]), F_,
                           Hg, Wg, ws_, c, 2 if r16 else self.rnd, _ptr(agg_ws), self.stream)
                continue
            self._meta = (4 * (2 * F_ * N * Dv + 2 * F_ * N * capi.SAB_SLOTS), 2 * F_ * N * 46 * Dv)
            self._call("turtle_sab_aggregate", _ptr(idx[b]), _ptr(wgt[b]), vf, N * Dv, _ptr(agg[b]), F_, Hg, Wg, ws_, c,
                 1 if t0 else 0, 2 if r16 else self.rnd, self.stream)
        if self.trace is not None:
            self.trace.setdefault(sa, []).append(dict(idx=idx.clone(), wgt=wgt.clone(), qn=qn.clone()))
        xs = self.ws.get("sab_xs16" if r16 else "sab_xs", B, F_, Pimg, c, dtype=dt16)
        self.conv1x1(_ptr(agg), c, c, sa + "project_out.weight", _ptr(xs), c, B * F_ * Pimg, c,
                     bias=self._w(sa + "project_out.bias"), round_out=True, a16=r16, o16=r16)
        ring.commit()
        k_out, v_out = ring.views()

        # --- router over the aligned history (T1:649-660) ---------------------------------------
        kv = self.ws.get("chm_kv16" if r16 else "chm_kv", B * F_ * Pimg, 2 * c, dtype=dt16)
        self.conv1x1(_ptr(xs), c, c, pre + "kv.weight", _ptr(kv), 2 * c, B * F_ * Pimg, 2 * c, bias=self._w(pre + "kv.bias"),
                     a16=r16, o16=r16)
        kvd = self.ws.get("chm_kvd16" if r16 else "chm_kvd", B, F_, Pimg, 2 * c, dtype=dt16)
        if r16:
            self._meta = (2 * B * F_ * Pimg * 4 * c, 2 * 9 * B * F_ * Pimg * 2 * c)
            self._call("turtle_dwconv3x3", _ptr(kv), 2 * c, _ptr(self._w(pre + "kv_dwconv.weight", "dw16")),
                       _ptr(self._w(pre + "kv_dwconv.bias")), _ptr(kvd), 2 * c, B * F_, H, W, 2 * c, 0, 0, 1, 2, self.stream)
        else:
            self.dwconv(_ptr(kv), 2 * c, pre + "kv_dwconv.weight", pre + "kv_dwconv.bias", _ptr(kvd), 2 * c, B * F_, H, W,
                        2 * c)
        ch = c // heads
        hist = []
        for b in range(B):
            segs = []
            for f in range(F_):
                base = (b * F_ + f) * Pimg * 2 * c
                segs.append(dict(k=_ptr(kvd, base), ldk=2 * c, khs=ch, v=_ptr(kvd, base + c), ldv=2 * c, vhs=ch,
                                 prenorm=False))
            hist.append(segs)
        self.channel_attn(pre + "ChanAttn.", xn, x, B, H, W, c, heads, hist_segs=hist, ln=ln, force16=r16)
        return k_out, v_out

    # ------------------------------------------------------------------------------------
    # block / level
    # ------------------------------------------------------------------------------------
    def _ffn_half(self, pre, blk, c) -> bool:
        if blk.FFW_type == "GFFW":
            return self.half_path(c, self._sd[pre + "ffn.project_in.weight"].shape[0] // 2)
        return self.half_path(c, 2 * c)

    def _attn_half(self, blk, c, lvl) -> bool:
        at = blk.attention_type
        return ((at == "ReducedAttn" and self.half_path(c, 2 * c)) or
                (at == "Channel" and self.half_path(c) and c // lvl.num_heads == 64))

    def _first_norm(self, pre, blk, c, lvl):
        """(prefix, consumer takes fp16) of the first LayerNorm a block applies to its input."""
        if blk.attention_type == "NoAttn":
            return pre + "norm2.", self._ffn_half(pre, blk, c)
        return pre + "norm1.", self._attn_half(blk, c, lvl)

    def block(self, pre, blk, x, B, H, W, c, lvl, k_in=None, v_in=None, next_norm=None):
        """next_norm: ``_first_norm`` of the block that reads x next (same map), or None.  Whenever the reader takes
        fp16 and the row fits one n-group, the GEMM that finishes a residual update also emits that LayerNorm."""
        P = B * H * W
        kc = vc = None
        at = blk.attention_type
        half = self._ffn_half(pre, blk, c)
        if at != "NoAttn":
            xn = self.layernorm(x, pre + "norm1.", c, P, half=self._attn_half(blk, c, lvl))
            a = pre + "attn."
            ln2 = pre + "norm2." if half and self.ln_fusable(c, pre + "norm2.") else None
            if at == "Channel":
                self.channel_attn(a, xn, x, B, H, W, c, lvl.num_heads, ln=ln2)
            elif at == "ReducedAttn":
                self.reduced_attn(a, xn, x, P, c, H, W, B, ln=ln2)
            elif at == "FHR":
                kc, vc = self.fhr(a, xn, x, B, H, W, c, lvl.num_heads, lvl.num_frames_tocache, k_in, v_in, ln=ln2)
            elif at == "CHM":
                kc, vc = self.chm(a, xn, x, B, H, W, c, lvl.num_heads, lvl.Scale_patchsize, lvl.num_frames_tocache,
                                  k_in, v_in, ln=ln2)
        xn = self.layernorm(x, pre + "norm2.", c, P, half=half)
        nl = None
        if next_norm is not None and next_norm[1] and xn.dtype == torch.float16 and self.ln_fusable(c, next_norm[0]):
            nl = next_norm[0]
        if blk.FFW_type == "GFFW":
            self.gated_ffw(pre + "ffn.", xn, x, P, c, H, W, B, ln=nl)
        else:
            self.plain_ffw(pre + "ffn.", xn, x, P, c, ln=nl)
        return kc, vc

    def level(self, name, x, B, H, W, k_in=None, v_in=None):
        lvl = getattr(self.model, name)
        blocks = list(lvl.transformer_blocks)
        n = len(blocks)
        kc = vc = None
        for i, blk in enumerate(blocks):
            last = i == n - 1
            nxt = None if last else self._first_norm(f"{name}.transformer_blocks.{i + 1}.", blocks[i + 1], lvl.dim, lvl)
            kc, vc = self.block(f"{name}.transformer_blocks.{i}.", blk, x, B, H, W, lvl.dim, lvl,
                                k_in if last else None, v_in if last else None, next_norm=nxt)
        return kc, vc

    def latent(self, x, B, H, W, k1, v1, k2, v2):
        lvl = self.model.latent
        blocks = list(lvl.transformer_blocks)
        n = len(blocks)
        out = [None] * 4
        for i, blk in enumerate(blocks):
            pre = f"latent.transformer_blocks.{i}."
            nxt = None if i == n - 1 else self._first_norm(f"latent.transformer_blocks.{i + 1}.", blocks[i + 1], lvl.dim, lvl)
            if i == 0:
                out[0], out[1] = self.block(pre, blk, x, B, H, W, lvl.dim, lvl, k1, v1, next_norm=nxt)
            elif i == n - 1:
                out[2], out[3] = self.block(pre, blk, x, B, H, W, lvl.dim, lvl, k2, v2, next_norm=nxt)
            else:
                self.block(pre, blk, x, B, H, W, lvl.dim, lvl, next_norm=nxt)
        return out

    def conv3x3(self, x, Cin, wname, out, ldo, B, H, W, Cout, store, round_out=False):
        if self.half_path(Cin) and Cin % 64 == 0:
            # tensor-core mode: an fp16 copy of the map feeds a kind::f16 implicit GEMM (twice the MMA rate of TF32,
            # half the A traffic of the 9 shifted tile loads)
            n = B * H * W * Cin
            x16 = self.ws.get("x16", n, dtype=torch.float16)
            self._meta = (6 * n, 0)
            self._call("turtle_cast_f16", _ptr(x), _ptr(x16), n, self.stream)
            self.gemm([(_ptr(x16), Cin)], Cin, self._w(wname, "conv316"), _ptr(out), ldo, B * H * W, Cout, im2col=1,
                      geom=(B, H, W), store=store, round_out=round_out, a16=True)
            return
        self.gemm([(_ptr(x), Cin)], Cin, self._w(wname, "conv3"), _ptr(out), ldo, B * H * W, Cout, im2col=1,
                  geom=(B, H, W), store=store, round_out=round_out)

    # ------------------------------------------------------------------------------------
    # whole frame
    # ------------------------------------------------------------------------------------
    # ------------------------------------------------------------------------------------
    # CUDA-graph replay of the steady-state frame (model.enable_cuda_graphs())
    #
    # A frame is ~520 dependent launches; at small frame sizes (256x256 crops, the 320x320 tiles of tiled inference)
    # the host cannot issue them as fast as the GPU retires them, and even at 1280x720 the per-launch gaps add up.
    # Once every history ring of the clip is full, the device addresses a frame touches depend only on the joint ring
    # state, which repeats every history.RING_PERIOD frames, so each state is captured once and replayed afterwards.
    # The key holds the ring identities, so independent histories (clips, tiles) get their own graphs.
    # ------------------------------------------------------------------------------------
    MAX_GRAPHS = 48      # default cap (8 histories x 6 states); a cached graph keeps its history rings alive, so the
                         # cap also bounds the HBM held for clips that have ended (oldest graphs are dropped first)

    @torch.no_grad()
    def _graph_forward(self, inp, k_cached, v_cached):
        if k_cached is None or v_cached is None:
            return None
        rings = []
        for k, v in zip(k_cached, v_cached):
            if k is None and v is None:
                continue
            r = resolve_ring(k, v)
            if r is None or r.count < r.keep:
                return None                       # foreign caches or a history still filling up: eager frame
            rings.append(r)
        if not rings or not self.packed or self._weights_fingerprint() != self._weights_fp:
            return None                           # weights not packed yet / updated in place since: eager frame repacks
        m = self.model
        key = (tuple(inp.shape), inp.device.index, m.precision, bool(getattr(m, "half_intermediates", True)),
               bool(getattr(m, "fuse_layernorm", True)), bool(getattr(m, "sab_front_half", True)),
               bool(getattr(m, "fuse_gffw", False)), bool(getattr(m, "gffw_tail", False)),
               bool(getattr(m, "sab_agg_tc", True)),
               tuple((r.serial, r.pos, r.shadow_ok() if hasattr(r, "shadow_ok") else None) for r in rings))
        ent = self.graphs.get(key)
        if ent is None:
            if key not in self._graph_seen:
                # first visit of this state: run it eagerly, so that every lazily built host-side object it needs (flag
                # tensors, tensor maps, workspace growth) exists before the capture of its second visit
                if len(self._graph_seen) > 4 * self.max_graphs:
                    self._graph_seen.clear()
                self._graph_seen.add(key)
                return None
            while len(self.graphs) >= self.max_graphs:
                self.graphs.pop(next(iter(self.graphs)))
            static_in = torch.empty_like(inp, memory_format=torch.contiguous_format)
            static_in.copy_(inp)
            n0 = capi.launch_count
            # host-side ring state before the capture: a capture that fails launched nothing, but ran the host protocol
            pre_state = [(r.pos, r.count, r._sig, r.epoch, r.v16_lo) for r in rings]
            g = torch.cuda.CUDAGraph()
            import gc
            gc_was = gc.isenabled()
            gc.collect()                     # dead cycles that own CUDA graphs / memory (an earlier model and its engine) go
            gc.disable()                     # NOW: a collection in the middle of a capture would destroy them there, and
            try:                             # freeing device memory while a stream captures invalidates the capture
                with torch.cuda.graph(g):    # private memory pool per graph: replay order is free, eviction frees it
                    out, ks, vs = self._forward_eager(static_in, k_cached, v_cached)
            except Exception as e:           # never let a failed capture cost a frame: restore the rings, run eagerly
                if gc_was:
                    gc.enable()
                for r, (pos, count, sig, ep, v16) in zip(rings, pre_state):
                    r.pos, r.count, r._sig, r.epoch, r.v16_lo = pos, count, sig, ep, v16
                capi.launch_count = n0
                self._capture_failures = getattr(self, "_capture_failures", 0) + 1
                self._graph_seen.discard(key)        # it gets another chance on a later visit
                import warnings
                warnings.warn(f"CUDA-graph capture of a frame failed ({type(e).__name__}: {str(e)[:120]}); frame runs eagerly")
                try:
                    torch.cuda.synchronize()
                except Exception:
                    pass
                if self._capture_failures > 8:
                    self.model.cuda_graphs = False   # something systematic: stop trying
                return None
            if gc_was:
                gc.enable()
            post = [(r.pos, r.count, r._sig, r.v16_lo) for r in rings]
            ent = self.graphs[key] = (g, static_in, out, ks, vs, rings, post, capi.launch_count - n0)
            self.graph_captures += 1
        else:
            g, static_in, out, ks, vs, rings, post, nl = ent
            static_in.copy_(inp)
            for r, (pos, count, sig, v16) in zip(rings, post):
                r.pos, r.count, r._sig, r.v16_lo = pos, count, sig, v16
            capi.launch_count += nl
        ent[0].replay()
        self.graph_replays += 1
        return ent[2].clone(), list(ent[3]), list(ent[4])

    @torch.no_grad()
    def forward(self, inp: torch.Tensor, k_cached=None, v_cached=None):
        if (getattr(self.model, "cuda_graphs", False) and not self.dry_run and self.profile is None
                and not getattr(self.model, "record_trace", False) and inp.is_cuda and inp.dim() == 5):
            r = self._graph_forward(inp.float(), k_cached, v_cached)
            if r is not None:
                return r
        return self._forward_eager(inp, k_cached, v_cached)

    @torch.no_grad()
    def _forward_eager(self, inp: torch.Tensor, k_cached=None, v_cached=None):
        m = self.model
        if not inp.is_cuda and not self.dry_run:
            raise RuntimeError("turtlevsr_b200 runs on CUDA tensors only (no CPU fallback)")
        if inp.dim() != 5 or inp.shape[1] != 2:
            raise ValueError("expected input of shape [B, 2, C, H, W]")
        self.device = inp.device
        first_param = next(m.parameters())
        if first_param.device != inp.device:
            raise RuntimeError(f"model is on {first_param.device}, input on {inp.device}")
        if self.ws is None or self.ws.device != inp.device:
            self.ws = _Workspace(inp.device, on_realloc=self.drop_graphs)
            self.invalidate()
        self._sd = dict(m.named_parameters())
        self._plist = list(self._sd.values())
        fp = self._weights_fingerprint()
        if fp != getattr(self, "_weights_fp", None):
            if self.packed:
                self.invalidate()
            self._weights_fp = fp
        self.mode = capi.TF32 if m.precision == "tf32" else capi.FP32
        self.rnd = 1 if self.mode == capi.TF32 else 0
        self.use_half = bool(getattr(m, "half_intermediates", True))
        self.fuse_ln = bool(getattr(m, "fuse_layernorm", True))
        self.sab_front_half = bool(getattr(m, "sab_front_half", True))
        self.fuse_gffw = (bool(getattr(m, "fuse_gffw", False)) or os.environ.get("TURTLE_FUSE_GFFW", "0") == "1")
        self.gffw_tail = (bool(getattr(m, "gffw_tail", False)) or os.environ.get("TURTLE_GFFW_TAIL", "0") == "1")
        self.gram_ctas = int(os.environ.get("TURTLE_GRAM_CTAS", "296"))      # pixel splits x heads of the Gram kernel
        self._fused = None
        self.trace = {} if getattr(m, "record_trace", False) else None
        inp = inp.float().contiguous()
        B, _, Cc, Hs, Ws = inp.shape
        up = 4 if m.variant == "super" else 1
        H, W = Hs * up, Ws * up
        Hp, Wp = H + (-H) % 32, W + (-W) % 32
        if k_cached is None:
            k_cached, v_cached = [None] * 8, [None] * 8
        dim = m.dim
        import contextlib
        with (contextlib.nullcontext() if self.dry_run else torch.cuda.device(inp.device)):
            self.stream = 0 if self.dry_run else torch.cuda.current_stream().cuda_stream
            ws = self.ws
            if m.use_both_input:
                Ci, src, bstride = 2 * Cc, _ptr(inp), 2 * Cc * Hs * Ws
            else:
                Ci, src, bstride = Cc, _ptr(inp, Cc * Hs * Ws), 2 * Cc * Hs * Ws
            img = ws.get("img", B, Hp, Wp, Ci)
            self._meta = (4 * B * Ci * (Hs * Ws + Hp * Wp), 0)
            self._call("turtle_pack_frame", src, bstride, _ptr(img), B, Ci, Hs, Ws, Hp, Wp, up, self.stream)

            ks: List[Optional[torch.Tensor]] = []
            vs: List[Optional[torch.Tensor]] = []
            e1 = ws.get("e1", B * Hp * Wp, dim)
            self._call("turtle_conv3x3_first", _ptr(img), _ptr(self._w("input_projection.weight")),
                 _ptr(self._w("input_projection.bias")), _ptr(e1), B, Hp, Wp, Ci, dim, self.stream)
            kc, vc = self.level("encoder_level1", e1, B, Hp, Wp, k_cached[0], v_cached[0]); ks.append(kc); vs.append(vc)
            H2, W2 = Hp // 2, Wp // 2
            e2 = ws.get("e2", B * H2 * W2, dim * 2)
            self.conv3x3(e1, dim, "down1_2.body.0.weight", e2, dim * 2, B, Hp, Wp, dim // 2, capi.STORE_UNSHUFFLE2)
            kc, vc = self.level("encoder_level2", e2, B, H2, W2, k_cached[1], v_cached[1]); ks.append(kc); vs.append(vc)
            H3, W3 = H2 // 2, W2 // 2
            e3 = ws.get("e3", B * H3 * W3, dim * 4)
            self.conv3x3(e2, dim * 2, "down2_3.body.0.weight", e3, dim * 4, B, H2, W2, dim, capi.STORE_UNSHUFFLE2)
            kc, vc = self.level("encoder_level3", e3, B, H3, W3, k_cached[2], v_cached[2]); ks.append(kc); vs.append(vc)
            H4, W4 = H3 // 2, W3 // 2
            x4 = ws.get("x4", B * H4 * W4, dim * 8)
            self.conv3x3(e3, dim * 4, "down3_4.body.0.weight", x4, dim * 8, B, H3, W3, dim * 2, capi.STORE_UNSHUFFLE2)
            k4, v4, k5, v5 = self.latent(x4, B, H4, W4, k_cached[3], v_cached[3], k_cached[4], v_cached[4])
            ks += [k4, k5]; vs += [v4, v5]

            def up_merge(xlow, Hl, Wl, Cl, upname, skip, redname, outname):
                # Upsample: 3x3 Cl->2Cl + PixelShuffle(2) => [2Hl,2Wl,Cl/2]; cat(skip) ; 1x1 Cl->Cl/2
                Ch = Cl // 2
                u = ws.get("up", B * 4 * Hl * Wl, Ch)
                self.conv3x3(xlow, Cl, upname + ".body.0.weight", u, Ch, B, Hl, Wl, 2 * Cl, capi.STORE_SHUFFLE2, round_out=True)
                d = ws.get(outname, B * 4 * Hl * Wl, Ch)
                self.gemm([(_ptr(u), Ch), (_ptr(skip), Ch)], Ch, self._w(redname + ".weight", "gemm"), _ptr(d), Ch,
                          B * 4 * Hl * Wl, Ch, bias=self._w(redname + ".bias"))
                return d

            d3 = up_merge(x4, H4, W4, dim * 8, "up4_3", e3, "reduce_chan_level3", "d3")
            kc, vc = self.level("decoder_level3", d3, B, H3, W3, k_cached[5], v_cached[5]); ks.append(kc); vs.append(vc)
            d2 = up_merge(d3, H3, W3, dim * 4, "up3_2", e2, "reduce_chan_level2", "d2")
            kc, vc = self.level("decoder_level2", d2, B, H2, W2, k_cached[6], v_cached[6]); ks.append(kc); vs.append(vc)
            d1 = up_merge(d2, H2, W2, dim * 2, "up2_1", e1, "reduce_chan_level1", "d1")
            kc, vc = self.level("decoder_level1", d1, B, Hp, Wp, k_cached[7], v_cached[7]); ks.append(kc); vs.append(vc)
            self.level("refinement", d1, B, Hp, Wp)

            out = torch.empty(B, m.out_channels, H, W, device=inp.device, dtype=torch.float32)
            self._call("turtle_conv3x3_last", _ptr(d1), _ptr(self._w("ending.weight")), _ptr(self._w("ending.bias")),
                 _ptr(img), Ci, Ci - Cc, _ptr(out), B, Hp, Wp, dim, m.out_channels, H, W, self.stream)
        if self.trace is not None:
            self.last_trace = self.trace
        return out, ks, vs
