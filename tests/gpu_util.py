"""Thin test-side wrappers that call the C ABI on torch CUDA tensors (NHWC fp32)."""
import ctypes as C

import torch

from turtlevsr_b200 import capi
from turtlevsr_b200.capi import GemmArgs, call


def stream():
    return torch.cuda.current_stream().cuda_stream


def nhwc(x):     # [B,C,H,W] -> [B,H,W,C] contiguous cuda
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def nchw(y):
    return y.permute(0, 3, 1, 2).contiguous().cpu()


def gemm(segs, segw, Wt, P, Cout, mode=0, bias=None, scale=None, act=0, res=None, im2col=0, geom=None, store=0,
         out=None, ldo=None):
    a = GemmArgs()
    a.mode, a.im2col, a.P, a.Cout, a.nseg, a.segw = mode, im2col, P, Cout, len(segs), segw
    if geom:
        a.B, a.H, a.W = geom
    for i, (t, off, ld) in enumerate(segs):
        a.A[i] = t.data_ptr() + 4 * off
        a.lda[i] = ld
    a.Wt = Wt.data_ptr()
    a.bias = None if bias is None else bias.data_ptr()
    a.scale = None if scale is None else scale.data_ptr()
    a.act = act
    if res is not None:
        a.res, a.ldres = res.data_ptr(), res.shape[-1]
    a.out, a.ldo, a.store = out.data_ptr(), ldo, store
    call("turtle_gemm", C.byref(a), stream())
    return out


_KEEP = []


def dp(t):
    """device pointer of a tensor moved to CUDA; the tensor is kept alive (a temporary's memory
    would be recycled by the caching allocator before the kernel reads it)."""
    t = t.cuda().contiguous()
    _KEEP.append(t)
    if len(_KEEP) > 256:
        torch.cuda.synchronize()
        del _KEEP[:128]
    return t.data_ptr()
