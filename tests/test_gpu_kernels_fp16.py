"""GPU parity of the fp16-storage variants of the C-ABI kernels (tensor-core mode intermediates).

Each kernel is checked against a torch fp64 restatement of the same op evaluated on the *same fp16-rounded
inputs*, so the tolerance only has to cover fp32 accumulation plus the final fp16 rounding (2^-11 relative)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from turtlevsr_b200 import capi  # noqa: E402
from turtlevsr_b200.capi import GemmArgs, call  # noqa: E402
from gpu_util import stream  # noqa: E402

H_EPS = 2.0 ** -11


def rnd(*s, seed=0):
    g = torch.Generator().manual_seed(seed + sum(s))
    return torch.randn(*s, generator=g)


def close16(got, want, extra=1e-6):
    """|got - want| <= one fp16 rounding of want (+ accumulation slack)."""
    err = (got.double() - want).abs()
    tol = want.abs() * (1.5 * H_EPS) + extra
    return bool((err <= tol).all()), float((err / tol).max())


@pytest.mark.parametrize("fuse", [0, 1, 2])
@pytest.mark.parametrize("B,C_,H,W", [(2, 64, 9, 14), (1, 320, 24, 40), (1, 192, 17, 33), (2, 128, 37, 300), (1, 64, 70, 129),
                                      (1, 640, 23, 160)])
def test_dwconv_fp16(fuse, B, C_, H, W):
    x = rnd(B, H, W, C_).half().cuda()                       # channels-last
    w = (rnd(C_, 1, 3, 3, seed=1) / 3).half()
    b = rnd(C_, seed=2)
    xd = x.cpu().double().permute(0, 3, 1, 2)
    y = F.conv2d(xd, w.double(), b.double(), padding=1, groups=C_)
    if fuse == 1:
        y = F.gelu(y)
    elif fuse == 2:
        a, g = y.chunk(2, 1)
        y = F.gelu(a) * g
    Co = y.shape[1]
    out = torch.full((B, H, W, Co), float("nan"), device="cuda", dtype=torch.float16)
    w9 = w.reshape(C_, 9).t().contiguous().cuda()
    bd = b.cuda()
    call("turtle_dwconv3x3", x.data_ptr(), C_, w9.data_ptr(), bd.data_ptr(), out.data_ptr(), Co, B, H, W,
         C_, fuse, 0, 1, 2, stream())
    ok, worst = close16(out.cpu().permute(0, 3, 1, 2), y, extra=2e-5)
    assert ok, worst


@pytest.mark.parametrize("fuse", [0, 1])
def test_dwconv_fp16_odd_block_count(fuse):
    """96 channels = three 32-channel blocks: the plain / GELU kernels cannot pair them up and take one block per item."""
    B, C_, H, W = 1, 96, 11, 19
    x = rnd(B, H, W, C_).half().cuda()
    w = (rnd(C_, 1, 3, 3, seed=1) / 3).half()
    b = rnd(C_, seed=2)
    y = F.conv2d(x.cpu().double().permute(0, 3, 1, 2), w.double(), b.double(), padding=1, groups=C_)
    if fuse == 1:
        y = F.gelu(y)
    out = torch.full((B, H, W, C_), float("nan"), device="cuda", dtype=torch.float16)
    w9, bd = w.reshape(C_, 9).t().contiguous().cuda(), b.cuda()          # (kept alive until the kernel has run)
    call("turtle_dwconv3x3", x.data_ptr(), C_, w9.data_ptr(), bd.data_ptr(), out.data_ptr(), C_, B, H, W, C_, fuse, 0, 1, 2,
         stream())
    ok, worst = close16(out.cpu().permute(0, 3, 1, 2), y, extra=2e-5)
    assert ok, worst


@pytest.mark.parametrize("C_", [64, 128, 256, 512])
def test_layernorm_fp16_out(C_):
    P = 777
    x = (rnd(P, C_) * 2 + 0.5).cuda()
    w, b = rnd(C_, seed=1).cuda(), rnd(C_, seed=2).cuda()
    xd = x.cpu().double()
    mu = xd.mean(-1, keepdim=True)
    var = ((xd - mu) ** 2).mean(-1, keepdim=True)
    want = (xd - mu) / (var + 1e-5).sqrt() * w.cpu().double() + b.cpu().double()
    y = torch.full((P, C_), float("nan"), device="cuda", dtype=torch.float16)
    call("turtle_layernorm", x.data_ptr(), C_, w.data_ptr(), b.data_ptr(), y.data_ptr(), C_, P, C_, 2, stream())
    ok, worst = close16(y.cpu(), want, extra=1e-5)
    assert ok, worst


def _gemm16(A, lda, segs_off, segw, Wt, P, Cout, out, ldo, o16, bias=None, scale=None, act=0, res=None):
    a = GemmArgs()
    a.mode, a.im2col, a.P, a.Cout, a.nseg, a.segw = capi.TF32, 0, P, Cout, len(segs_off), segw
    for i, off in enumerate(segs_off):
        a.A[i] = A.data_ptr() + 2 * off
        a.lda[i] = lda
    a.Wt = Wt.data_ptr()
    a.bias = None if bias is None else bias.data_ptr()
    a.scale = None if scale is None else scale.data_ptr()
    a.act = act
    if res is not None:
        a.res, a.ldres = res.data_ptr(), res.shape[-1]
    a.out, a.ldo, a.store = out.data_ptr(), ldo, capi.STORE_PLAIN
    a.a_dtype, a.out_dtype = 1, 1 if o16 else 0
    call("turtle_gemm", C.byref(a), stream())


@pytest.mark.parametrize("Cin,Cout,P", [(64, 320, 1000), (256, 768, 333), (512, 2560, 300)])
def test_gemm_fp16_in_fp16_out(Cin, Cout, P):
    A = rnd(P, Cin).half().cuda()
    Wt = (rnd(Cout, Cin, seed=3) / Cin ** 0.5).half().cuda()
    bias = rnd(Cout, seed=4).cuda()
    out = torch.full((P, Cout), float("nan"), device="cuda", dtype=torch.float16)
    _gemm16(A, Cin, [0], Cin, Wt, P, Cout, out, Cout, True, bias=bias, act=capi.ACT_GELU)
    want = F.gelu(A.cpu().double() @ Wt.cpu().double().t() + bias.cpu().double())
    ok, worst = close16(out.cpu(), want, extra=2e-5)
    assert ok, worst


@pytest.mark.parametrize("Cin,Cout,P", [(160, 64, 777), (1280, 512, 300), (128, 64, 1000)])
def test_gemm_fp16_in_fp32_out_residual(Cin, Cout, P):
    A = rnd(P, Cin).half().cuda()
    Wt = (rnd(Cout, Cin, seed=3) / Cin ** 0.5).half().cuda()
    bias, scale = rnd(Cout, seed=4).cuda(), rnd(Cout, seed=5).cuda()
    x = rnd(P, Cout, seed=6).cuda()
    want = (A.cpu().double() @ Wt.cpu().double().t() + bias.cpu().double()) * scale.cpu().double() + x.cpu().double()
    _gemm16(A, Cin, [0], Cin, Wt, P, Cout, x, Cout, False, bias=bias, scale=scale, res=x)
    assert (x.cpu().double() - want).abs().max() < 1e-5 * want.abs().max()        # fp32 accumulation over K


def test_gemm_fp16_segments():
    """Folded channel-attention apply: K-concatenated fp16 value blocks (one per head), fp16 folded weight."""
    P, ch, heads = 500, 64, 4
    c = heads * ch
    qkv = rnd(P, 3 * c).half().cuda()
    M = (rnd(c, c, seed=9) / 16).half().cuda()
    x = rnd(P, c, seed=10).cuda()
    want = x.cpu().double() + qkv.cpu().double()[:, 2 * c:] @ M.cpu().double().t()
    _gemm16(qkv, 3 * c, [2 * c + h * ch for h in range(heads)], ch, M, P, c, x, c, False, res=x)
    assert (x.cpu().double() - want).abs().max() < 1e-5 * want.abs().max()


@pytest.mark.parametrize("P,heads,nsplit", [(640, 1, 1), (700, 2, 5), (4096, 4, 7), (130, 8, 3)])
def test_chan_gram_fp16(P, heads, nsplit):
    c = heads * 64
    x = rnd(P, 3 * c).half().cuda()
    g = torch.full((nsplit, heads, 64, 64), float("nan"), device="cuda")
    sqq = torch.full((nsplit, c), float("nan"), device="cuda")
    sqk = torch.full((nsplit, c), float("nan"), device="cuda")
    call("turtle_chan_gram", x.data_ptr(), 3 * c, 64, x.data_ptr() + 2 * c, 3 * c, 64, P, heads, 64, nsplit,
         g.data_ptr(), sqq.data_ptr(), sqk.data_ptr(), 2, stream())
    xd = x.cpu().double()
    q, k = xd[:, :c].reshape(P, heads, 64), xd[:, c:2 * c].reshape(P, heads, 64)
    want = torch.einsum("phi,phj->hij", q, k)
    assert (g.sum(0).cpu().double() - want).abs().max() < 1e-5 * P ** 0.5 * 8
    assert (sqq.sum(0).cpu().double() - (xd[:, :c] ** 2).sum(0)).abs().max() < 1e-5 * P
    assert (sqk.sum(0).cpu().double() - (xd[:, c:2 * c] ** 2).sum(0)).abs().max() < 1e-5 * P


@pytest.mark.parametrize("a16", [True, False])
@pytest.mark.parametrize("Cin,C_,P", [(160, 64, 777), (320, 128, 1000), (640, 256, 300), (64, 64, 130)])
def test_gemm_fused_layernorm(Cin, C_, P, a16):
    """x += A W^T (+bias)(*scale); ln_out = fp16(LayerNorm(x) * w + b) produced by the same launch."""
    A = rnd(P, Cin)
    Wt = rnd(C_, Cin, seed=3) / Cin ** 0.5
    if a16:
        A, Wt = A.half(), Wt.half()
    A, Wt = A.cuda(), Wt.cuda()
    bias, scale = rnd(C_, seed=4).cuda(), rnd(C_, seed=5).cuda()
    lw, lb = (rnd(C_, seed=7) * 0.5 + 1).cuda(), rnd(C_, seed=8).cuda()
    x = (rnd(P, C_, seed=6) * 2 + 0.7).cuda()
    xn = torch.full((P, C_), float("nan"), device="cuda", dtype=torch.float16)
    want_x = (A.cpu().double() @ Wt.cpu().double().t() + bias.cpu().double()) * scale.cpu().double() + x.cpu().double()
    a = GemmArgs()
    a.mode, a.im2col, a.P, a.Cout, a.nseg, a.segw = capi.TF32, 0, P, C_, 1, Cin
    a.A[0], a.lda[0] = A.data_ptr(), Cin
    a.Wt, a.bias, a.scale = Wt.data_ptr(), bias.data_ptr(), scale.data_ptr()
    a.res, a.ldres, a.out, a.ldo, a.store = x.data_ptr(), C_, x.data_ptr(), C_, capi.STORE_PLAIN
    a.a_dtype = 1 if a16 else 0
    a.ln_out, a.ld_ln, a.ln_w, a.ln_b = xn.data_ptr(), C_, lw.data_ptr(), lb.data_ptr()
    call("turtle_gemm", C.byref(a), stream())
    got_x = x.cpu().double()
    tol = (1e-5 if a16 else 3e-3) * want_x.abs().max()
    assert (got_x - want_x).abs().max() < tol
    # the normalisation is checked against the rows the kernel actually stored
    mu = got_x.mean(-1, keepdim=True)
    var = ((got_x - mu) ** 2).mean(-1, keepdim=True)
    want_n = (got_x - mu) / (var + 1e-5).sqrt() * lw.cpu().double() + lb.cpu().double()
    ok, worst = close16(xn.cpu(), want_n, extra=2e-5)
    assert ok, worst


@pytest.mark.parametrize("bias", [False, True])
@pytest.mark.parametrize("ws,D", [(4, 64), (8, 128), (16, 32), (4, 256), (4, 512), (16, 128)])
def test_sab_window_reduce_fp16_map(ws, D, bias):
    from oracle import turtle_oracle as orc
    B, H, W = 2, 32, 48
    t = rnd(B, H, W, D).half().cuda()
    w = rnd(D, 1, ws, ws, seed=1)
    td = t.cpu().float().permute(0, 3, 1, 2)
    b = rnd(D, seed=2).cuda() if bias else None
    want = orc.l2norm_rows(F.conv2d(td, w, None if b is None else b.cpu(), stride=ws, padding=1,
                                    groups=D).flatten(2).transpose(1, 2))
    N = (H // ws) * (W // ws)
    out = torch.full((B, N, D), float("nan"), device="cuda")
    wk = w.reshape(D, -1).t().contiguous().cuda()
    call("turtle_sab_window_reduce_h16", t.data_ptr(), D, wk.data_ptr(), b.data_ptr() if bias else None, out.data_ptr(),
         N * D, B, H, W, D, ws, stream())
    assert (out.cpu() - want).abs().max() < 2e-5


@pytest.mark.parametrize("store", [capi.STORE_UNSHUFFLE2, capi.STORE_SHUFFLE2, capi.STORE_PLAIN])
def test_conv3x3_im2col_fp16_operands(store):
    """Dense 3x3 (Down/Upsample, T1:136-154) as a kind::f16 implicit GEMM on an fp16 copy of the map."""
    B, Cin, H, W = 2, 64, 12, 20
    Cout = 32 if store == capi.STORE_UNSHUFFLE2 else 128
    x = rnd(B, H, W, Cin).cuda()
    x16 = torch.full((B, H, W, Cin), float("nan"), device="cuda", dtype=torch.float16)
    call("turtle_cast_f16", x.data_ptr(), x16.data_ptr(), x.numel(), stream())
    assert torch.equal(x16, x.half())
    w = (rnd(Cout, Cin, 3, 3, seed=1) / (9 * Cin) ** 0.5).half()
    y = F.conv2d(x16.cpu().double().permute(0, 3, 1, 2), w.double(), padding=1)
    want = F.pixel_unshuffle(y, 2) if store == capi.STORE_UNSHUFFLE2 else F.pixel_shuffle(y, 2) if store == capi.STORE_SHUFFLE2 else y
    Bo, Co, Ho, Wo = want.shape
    out = torch.full((Bo, Ho, Wo, Co), float("nan"), device="cuda")
    wp = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().cuda()
    a = GemmArgs()
    a.mode, a.im2col, a.P, a.Cout, a.nseg, a.segw = capi.TF32, 1, B * H * W, Cout, 1, Cin
    a.B, a.H, a.W = B, H, W
    a.A[0], a.lda[0] = x16.data_ptr(), Cin
    a.Wt, a.out, a.ldo, a.store, a.a_dtype = wp.data_ptr(), out.data_ptr(), Co, store, 1
    call("turtle_gemm", C.byref(a), stream())
    assert (out.cpu().double().permute(0, 3, 1, 2) - want).abs().max() < 1e-5 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize("a16", [True, False])
@pytest.mark.parametrize("K,Cout,rows,B,ln", [(128, 128, 256, 3, True), (256, 256, 512, 2, False), (64, 64, 128, 5, True),
                                              (512, 256, 1024, 2, True)])
def test_gemm_per_batch_weights_equals_per_image_launches(K, Cout, rows, B, ln, a16):
    """TurtleGemmArgs.w_batches: one launch with a weight matrix per image (the folded channel-attention apply of B
    tiles) == B launches, bit for bit; shapes that do not tile into 128-row blocks per image are refused (ENOTSUP)."""
    P = B * rows
    dt = torch.float16 if a16 else torch.float32
    A = (rnd(P, K) * 0.5).to(dt).cuda()
    Wt = (rnd(B, Cout, K, seed=1) / K ** 0.5).to(dt).cuda()
    x0 = rnd(P, Cout, seed=2).cuda()
    lw, lb = (torch.rand(Cout) + 0.5).cuda(), rnd(Cout, seed=3).cuda()

    def run(out, lnout, b0, nb):
        a = GemmArgs()
        a.mode, a.P, a.Cout, a.nseg, a.segw = capi.TF32, nb * rows, Cout, 1, K
        a.A[0], a.lda[0] = A.data_ptr() + b0 * rows * K * A.element_size(), K
        a.Wt = Wt.data_ptr() + b0 * Cout * K * Wt.element_size()
        a.res = a.out = out.data_ptr() + b0 * rows * Cout * 4
        a.ldres = a.ldo = Cout
        a.a_dtype = 1 if a16 else 0
        if ln:
            a.ln_out, a.ld_ln, a.ln_w, a.ln_b = lnout.data_ptr() + b0 * rows * Cout * 2, Cout, lw.data_ptr(), lb.data_ptr()
        if nb > 1:
            a.w_batches, a.w_bstride, a.rows_per_batch = nb, Cout * K, rows
        call("turtle_gemm", C.byref(a), stream())

    one, many = x0.clone(), x0.clone()
    ln_one = torch.zeros(P, Cout, device="cuda", dtype=torch.float16)
    ln_many = torch.zeros_like(ln_one)
    run(one, ln_one, 0, B)
    for b in range(B):
        run(many, ln_many, b, 1)
    torch.cuda.synchronize()
    assert torch.equal(one, many)
    if ln:
        assert torch.equal(ln_one, ln_many)
    # reference: per-image matmul on the rounded operands
    want = x0.double() + torch.einsum("bpk,bok->bpo", A.double().view(B, rows, K).cpu().cuda(), Wt.double()).reshape(P, Cout)
    assert (one.double() - want).abs().max().item() < (2e-3 if not a16 else 1e-3) * max(1.0, want.abs().max().item())
    # rows per image not a multiple of the 128-row tile: refused, the caller launches per image
    a = GemmArgs()
    a.mode, a.P, a.Cout, a.nseg, a.segw = capi.TF32, 2 * 200, Cout, 1, K
    a.A[0], a.lda[0], a.Wt = A.data_ptr(), K, Wt.data_ptr()
    a.out, a.ldo, a.a_dtype = one.data_ptr(), Cout, 1 if a16 else 0
    a.w_batches, a.w_bstride, a.rows_per_batch = 2, Cout * K, 200
    assert capi.load().turtle_gemm(C.byref(a), stream()) == capi.ENOTSUP
