"""GPU: the fused GatedFeedForward kernel (csrc/gffw_fused.cu, T1:159-178) against (a) an fp32 torch restatement of the
oracle's gated_ffw on the same fp16-rounded operands and (b) the unfused C-ABI schedule it replaces
(turtle_gemm -> turtle_dwconv3x3(fuse=2) -> turtle_gemm).  Tolerance: the hidden map and the gated map are stored in
fp16 on both paths (11-bit significands), accumulation is fp32: 3e-3 relative to the output scale, written here."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from turtlevsr_b200 import capi  # noqa: E402
from turtlevsr_b200.capi import GemmArgs, call  # noqa: E402


def stream():
    return torch.cuda.current_stream().cuda_stream


def pack_taps(wdw, hid):
    return wdw.reshape(2, hid // 32, 32, 9).permute(1, 0, 3, 2).contiguous().half()


def reference(xn16, w_in16, wdw16, w_out16, x, B, H, W, c, hid):
    """oracle.gated_ffw (T1:173-178) in fp32 on the fp16-rounded operands, with the two fp16 roundings of the
    intermediates the tensor-core mode stores."""
    xn = xn16.float().view(B, H, W, c).permute(0, 3, 1, 2)
    t = F.conv2d(xn, w_in16.float().view(2 * hid, c, 1, 1)).half().float()
    u = F.conv2d(t, wdw16.float().view(2 * hid, 1, 3, 3), padding=1, groups=2 * hid)
    g = (F.gelu(u[:, :hid]) * u[:, hid:]).half().float()
    y = F.conv2d(g, w_out16.float().view(c, hid, 1, 1))
    return x + y.permute(0, 2, 3, 1).reshape(B * H * W, c)


def gemm16(A16, K, W16, out, ldo, P, Cout, o16, res=None):
    a = GemmArgs()
    a.mode, a.P, a.Cout, a.nseg, a.segw = capi.TF32, P, Cout, 1, K
    a.A[0], a.lda[0] = A16.data_ptr(), K
    a.Wt = W16.data_ptr()
    a.out, a.ldo = out.data_ptr(), ldo
    a.a_dtype, a.out_dtype = 1, 1 if o16 else 0
    if res is not None:
        a.res, a.ldres = res.data_ptr(), Cout
    call("turtle_gemm", C.byref(a), stream())


@pytest.mark.parametrize("c,hid,B,H,W,ln", [
    (64, 160, 1, 32, 48, True), (64, 160, 2, 24, 40, False), (128, 320, 1, 16, 32, True), (256, 640, 1, 24, 24, True),
    (64, 32, 1, 8, 16, False), (64, 64, 1, 10, 18, True), (128, 320, 2, 9, 21, True), (256, 640, 1, 46, 80, False),
])
def test_fused_gffw_matches_reference_and_unfused_path(c, hid, B, H, W, ln):
    g = torch.Generator().manual_seed(c * 1000 + hid + H)
    P = B * H * W
    xn16 = torch.randn(P, c, generator=g).half().cuda()
    w_in16 = (torch.randn(2 * hid, c, generator=g) / c ** 0.5).half().cuda()
    wdw16 = (torch.randn(2 * hid, 9, generator=g) / 3).half().cuda()
    w_out16 = (torch.randn(c, hid, generator=g) / hid ** 0.5).half().cuda()
    x0 = torch.randn(P, c, generator=g).cuda()
    ln_w, ln_b = (torch.rand(c, generator=g) + 0.5).cuda(), torch.randn(c, generator=g).cuda()
    want = reference(xn16, w_in16, wdw16, w_out16, x0, B, H, W, c, hid)

    x = x0.clone()
    ln_out = torch.full((P, c), float("nan"), device="cuda", dtype=torch.float16) if ln else None
    call("turtle_gffw_fused", xn16.data_ptr(), w_in16.data_ptr(), pack_taps(wdw16.float(), hid).cuda().data_ptr(),
         w_out16.data_ptr(), x.data_ptr(), ln_out.data_ptr() if ln else None, ln_w.data_ptr() if ln else None,
         ln_b.data_ptr() if ln else None, B, H, W, c, hid, stream())
    torch.cuda.synchronize()
    scale = want.abs().max().item()
    err = (x - want).abs().max().item()
    print(f"fused GFFW c={c} hid={hid} {B}x{H}x{W}: max|d| vs fp32 restatement {err:.3e} (output scale {scale:.2f})")
    assert err <= 3e-3 * max(1.0, scale)
    if ln:
        mu = x.mean(1, keepdim=True)
        var = x.var(1, unbiased=False, keepdim=True)
        want_ln = (x - mu) / torch.sqrt(var + 1e-5) * ln_w + ln_b
        assert (ln_out.float() - want_ln).abs().max().item() <= 4e-3 * max(1.0, want_ln.abs().max().item())

    # the unfused schedule on the same operands (only where its kernels accept the shape)
    if hid % 32 == 0 and (2 * hid) % 64 == 0:
        t16 = torch.empty(P, 2 * hid, device="cuda", dtype=torch.float16)
        gemm16(xn16, c, w_in16, t16, 2 * hid, P, 2 * hid, True)
        g16 = torch.empty(P, hid, device="cuda", dtype=torch.float16)
        taps = wdw16.float().t().contiguous().half()        # [9, 2*hid] tap-major, as engine._w(..., "dw16")
        call("turtle_dwconv3x3", t16.data_ptr(), 2 * hid, taps.data_ptr(), None, g16.data_ptr(), hid, B, H, W, 2 * hid, 2, 0,
             1, 2, stream())
        x2 = x0.clone()
        gemm16(g16, hid, w_out16, x2, c, P, c, False, res=x2)
        torch.cuda.synchronize()
        d = (x - x2).abs().max().item()
        print(f"   fused vs unfused schedule: max|d| {d:.3e}")
        assert d <= 3e-3 * max(1.0, scale)
        # the second half on its own (turtle_gffw_tail): depthwise + gate feeding project_out, from the same hidden map
        if hid * 2 <= 5 * c:
            x3 = x0.clone()
            ln3 = torch.full((P, c), float("nan"), device="cuda", dtype=torch.float16) if ln else None
            call("turtle_gffw_tail", t16.data_ptr(), pack_taps(wdw16.float(), hid).cuda().data_ptr(), w_out16.data_ptr(),
                 x3.data_ptr(), ln3.data_ptr() if ln else None, ln_w.data_ptr() if ln else None,
                 ln_b.data_ptr() if ln else None, B, H, W, c, hid, stream())
            torch.cuda.synchronize()
            d3 = (x3 - x2).abs().max().item()
            print(f"   gffw_tail vs unfused schedule: max|d| {d3:.3e}")
            assert d3 <= 3e-3 * max(1.0, scale)
            if ln:
                assert (ln3.float() - ln_out.float()).abs().max().item() <= 8e-3 * max(1.0, ln_out.float().abs().max().item())


def test_fused_gffw_rejects_unsupported_shapes():
    lib = capi.load()
    z = torch.zeros(64, device="cuda")
    p = z.data_ptr()
    assert lib.turtle_gffw_fused(p, p, p, p, p, None, None, None, 1, 8, 8, 512, 1280, None) == capi.ENOTSUP
    assert lib.turtle_gffw_fused(p, p, p, p, p, None, None, None, 1, 8, 8, 64, 100, None) == capi.ENOTSUP
    assert lib.turtle_gffw_fused(None, p, p, p, p, None, None, None, 1, 8, 8, 64, 160, None) == -1
