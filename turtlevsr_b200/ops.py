"""The torch custom-op layer over the C ABI: one ``torch.ops.turtle_b200.<name>`` per kernel-launching entry
point of include/turtle_b200.h.

The operator schemas are generated from the C prototypes in the header itself, so the two cannot drift apart:

  * a device pointer ``const float *x`` becomes ``Tensor? x, int x_off`` -- the tensor that owns the memory plus a
    BYTE offset into it (how the engine addresses channel slices, ring slots and batch elements of wider buffers);
    a non-const pointer is declared mutated (``Tensor(a!)? y``), which is what tells ``torch.compile`` /
    functionalisation that the op writes it;
  * ``int`` / ``int64_t`` / ``long long`` -> ``int``, ``float`` -> ``float``; host arrays (``const int *y0``) -> ``int[]``;
  * ``void *stream`` disappears: the implementation launches on ``torch.cuda.current_stream()``;
  * ``turtle_gemm`` takes its ``TurtleGemmArgs`` unpacked (segment tensors as a ``Tensor[]``).

Every op returns nothing (the kernels write caller-owned buffers), has a CUDA implementation that converts its
arguments back to raw pointers and calls the C ABI (``capi.call``: non-zero return codes raise), and a fake
(meta) implementation that does nothing, so graphs containing these ops can be traced without a GPU.  There is no
CPU implementation: calling an op on CPU tensors fails in the dispatcher.

``FrameEngine`` launches every kernel through these ops (``launch`` / ``launch_gemm`` below).  ``TURTLE_OPS=direct``
bypasses the dispatcher and calls ctypes directly (A/B measurements of the dispatch overhead).
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List, Optional, Sequence

import torch

from . import capi

NAMESPACE = "turtle_b200"
_HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "turtle_b200.h")
DIRECT = os.environ.get("TURTLE_OPS", "") == "direct"


class DevPtr(int):
    """A device address that remembers the tensor it points into (``base``).  It *is* an int, so ctypes takes it as
    a ``void *``; the op layer turns it back into (tensor, byte offset)."""

    def __new__(cls, base: torch.Tensor, byte_off: int = 0):
        self = super().__new__(cls, base.data_ptr() + byte_off)
        self.base = base
        return self

    def __add__(self, byte_off):                     # address arithmetic keeps the owner
        return DevPtr(self.base, int(self) - self.base.data_ptr() + int(byte_off))

    __radd__ = __add__


def devptr(t: Optional[torch.Tensor], elem_off: int = 0) -> Optional[DevPtr]:
    return None if t is None else DevPtr(t, t.element_size() * elem_off)


# ----------------------------------------------------------------------------------------------------------------
# schema generation from the header
# ----------------------------------------------------------------------------------------------------------------
class _Param:
    __slots__ = ("name", "kind", "mutated")          # kind: ptr | int | float | ints | stream

    def __init__(self, name, kind, mutated=False):
        self.name, self.kind, self.mutated = name, kind, mutated


def _parse_header() -> Dict[str, List[_Param]]:
    src = re.sub(r"/\*.*?\*/", "", open(_HEADER).read(), flags=re.S)
    protos: Dict[str, List[_Param]] = {}
    for m in re.finditer(r"\bint\s+(turtle_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        if "void *stream" not in args or name == "turtle_gemm":
            continue                                  # queries (abi version, workspace sizes) launch nothing
        params = []
        for a in args.split(","):
            a = a.strip()
            pname = re.search(r"([A-Za-z_][A-Za-z0-9_]*)$", a).group(1)
            ctype = a[: -len(pname)].strip()
            if pname == "stream":
                params.append(_Param(pname, "stream"))
            elif "*" in ctype:
                if re.match(r"const int\b", ctype):
                    params.append(_Param(pname, "ints"))          # host array (tile origins)
                else:
                    params.append(_Param(pname, "ptr", mutated=not ctype.startswith("const")))
            elif ctype in ("float",):
                params.append(_Param(pname, "float"))
            else:
                params.append(_Param(pname, "int"))
        protos[name] = params
    return protos


def _schema(params: Sequence[_Param]) -> str:
    out, alias = [], iter("abcdefghijklmnop")
    for p in params:
        if p.kind == "ptr":
            out.append((f"Tensor({next(alias)}!)? " if p.mutated else "Tensor? ") + p.name)
            out.append(f"int {p.name}_off")
        elif p.kind == "int":
            out.append(f"int {p.name}")
        elif p.kind == "float":
            out.append(f"float {p.name}")
        elif p.kind == "ints":
            out.append(f"int[] {p.name}")
    return "(" + ", ".join(out) + ") -> ()"


_GEMM_SCHEMA = (
    "(int mode, int im2col, int P, int B, int H, int W, int Cout, int segw, Tensor[] A, int[] A_off, int[] lda, "
    "Tensor Wt, int Wt_off, Tensor? bias, int bias_off, Tensor? scale, int scale_off, int act, Tensor? res, int res_off, "
    "int ldres, Tensor(a!) out, int out_off, int ldo, int store, int round_out, int a_dtype, int out_dtype, "
    "Tensor(b!)? ln_out, int ln_out_off, int ld_ln, Tensor? ln_w, int ln_w_off, Tensor? ln_b, int ln_b_off, "
    "int w_batches, int w_bstride, int rows_per_batch) -> ()")

PROTOS = _parse_header()
SCHEMAS: Dict[str, str] = {n[len("turtle_"):]: _schema(p) for n, p in PROTOS.items()}
SCHEMAS["gemm"] = _GEMM_SCHEMA

_lib = torch.library.Library(NAMESPACE, "FRAGMENT")


def _addr(t: Optional[torch.Tensor], off: int):
    return None if t is None else t.data_ptr() + off


def _current_stream(*tensors) -> int:
    for t in tensors:
        if isinstance(t, torch.Tensor):
            return torch.cuda.current_stream(t.device).cuda_stream
    return torch.cuda.current_stream().cuda_stream


def _make_impl(cname: str, params: Sequence[_Param]):
    def impl(*args):
        cargs, i, first = [], 0, None
        for p in params:
            if p.kind == "ptr":
                t, off = args[i], args[i + 1]
                i += 2
                if first is None and t is not None:
                    first = t
                cargs.append(_addr(t, off))
            elif p.kind == "ints":
                v = args[i]
                i += 1
                cargs.append((C.c_int32 * len(v))(*v))
            elif p.kind == "stream":
                cargs.append(_current_stream(first))
            else:
                cargs.append(args[i])
                i += 1
        capi.call(cname, *cargs)
    return impl


def _fake(*args, **kwargs):
    return None


def _gemm_impl(mode, im2col, P, B, H, W, Cout, segw, A, A_off, lda, Wt, Wt_off, bias, bias_off, scale, scale_off, act, res,
               res_off, ldres, out, out_off, ldo, store, round_out, a_dtype, out_dtype, ln_out, ln_out_off, ld_ln, ln_w,
               ln_w_off, ln_b, ln_b_off, w_batches, w_bstride, rows_per_batch):
    a = capi.GemmArgs()
    a.mode, a.im2col, a.P, a.B, a.H, a.W, a.Cout, a.nseg, a.segw = mode, im2col, P, B, H, W, Cout, len(A), segw
    for i, (t, o, ld) in enumerate(zip(A, A_off, lda)):
        a.A[i] = t.data_ptr() + o
        a.lda[i] = ld
    a.Wt, a.bias, a.scale, a.act = _addr(Wt, Wt_off), _addr(bias, bias_off), _addr(scale, scale_off), act
    a.res, a.ldres, a.out, a.ldo, a.store = _addr(res, res_off), ldres, _addr(out, out_off), ldo, store
    a.round_out, a.a_dtype, a.out_dtype = round_out, a_dtype, out_dtype
    a.ln_out, a.ld_ln, a.ln_w, a.ln_b = _addr(ln_out, ln_out_off), ld_ln, _addr(ln_w, ln_w_off), _addr(ln_b, ln_b_off)
    a.w_batches, a.w_bstride, a.rows_per_batch = w_batches, w_bstride, rows_per_batch
    capi.call("turtle_gemm", C.byref(a), _current_stream(out))


for _name, _sch in SCHEMAS.items():
    _lib.define(_name + _sch)
    if _name == "gemm":
        _lib.impl(_name, _gemm_impl, "CUDA")
    else:
        _lib.impl(_name, _make_impl("turtle_" + _name, PROTOS["turtle_" + _name]), "CUDA")
    torch.library.register_fake(f"{NAMESPACE}::{_name}", _fake, lib=_lib)

_OPS = {n: getattr(getattr(torch.ops, NAMESPACE), n) for n in SCHEMAS}


# ----------------------------------------------------------------------------------------------------------------
# engine-side entry: the same positional arguments as the C prototype (pointers as DevPtr / None)
# ----------------------------------------------------------------------------------------------------------------
def _split(p):
    if p is None:
        return None, 0
    if isinstance(p, DevPtr):
        return p.base, int(p) - p.base.data_ptr()
    raise TypeError("device pointers passed to the op layer must come from ops.devptr() (got a bare address)")


def launch(cname: str, *cargs) -> None:
    """``launch("turtle_layernorm", x_ptr, ldx, ...)`` with the C prototype's arguments (stream included, ignored)."""
    if DIRECT:
        capi.call(cname, *cargs)
        return
    params = PROTOS[cname]
    args = []
    for p, v in zip(params, cargs):
        if p.kind == "ptr":
            args.extend(_split(v))
        elif p.kind == "stream":
            continue
        elif p.kind == "ints":
            args.append(list(v))
        else:
            args.append(v)
    _OPS[cname[len("turtle_"):]](*args)


def launch_gemm(a: "capi.GemmArgs", ptrs: dict, stream) -> None:
    """``a``: the filled TurtleGemmArgs mirror; ``ptrs``: the DevPtr objects its pointer fields were set from."""
    if DIRECT:
        capi.call("turtle_gemm", C.byref(a), stream)
        return
    A = [_split(p) for p in ptrs["A"]]
    Wt, bias, scale, res = _split(ptrs["Wt"]), _split(ptrs.get("bias")), _split(ptrs.get("scale")), _split(ptrs.get("res"))
    out = _split(ptrs["out"])
    has_ln = bool(a.ln_out)
    ln_out = _split(ptrs.get("ln_out")) if has_ln else (None, 0)
    ln_w = _split(ptrs.get("ln_w")) if has_ln else (None, 0)
    ln_b = _split(ptrs.get("ln_b")) if has_ln else (None, 0)
    _OPS["gemm"](a.mode, a.im2col, a.P, a.B, a.H, a.W, a.Cout, a.segw, [t for t, _ in A], [o for _, o in A],
                 [a.lda[i] for i in range(a.nseg)], Wt[0], Wt[1], bias[0], bias[1], scale[0], scale[1], a.act, res[0],
                 res[1], a.ldres, out[0], out[1], a.ldo, a.store, a.round_out, a.a_dtype, a.out_dtype, ln_out[0],
                 ln_out[1], a.ld_ln, ln_w[0], ln_w[1], ln_b[0], ln_b[1], a.w_batches, a.w_bstride, a.rows_per_batch)
