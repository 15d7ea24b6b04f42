"""GPU parity of each C-ABI kernel against plain torch fp32 (CPU) restatements of the same op."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import turtle_oracle as orc  # noqa: E402
from turtlevsr_b200 import capi  # noqa: E402
from turtlevsr_b200.capi import call  # noqa: E402
from gpu_util import dp, gemm, nchw, nhwc, stream  # noqa: E402

TOL = 2e-5
TF32_REL = 3e-3      # tcgen05 kind::tf32 keeps 10 mantissa bits of each operand (K up to 512 here)


def rnd(*s, seed=0):
    g = torch.Generator().manual_seed(seed + sum(s))
    return torch.randn(*s, generator=g)


@pytest.mark.parametrize("C_", [64, 128, 512, 24])
def test_layernorm(C_):
    x = rnd(2, C_, 12, 20) * 2 + 0.5
    w, b = rnd(C_, seed=1), rnd(C_, seed=2)
    want = orc.channel_layernorm(x, w, b)
    xd = nhwc(x)
    y = torch.empty_like(xd)
    call("turtle_layernorm", xd.data_ptr(), C_, dp(w), dp(b), y.data_ptr(), C_,
         2 * 12 * 20, C_, 0, stream())
    assert (nchw(y) - want).abs().max() < TOL


@pytest.mark.parametrize("mode", [capi.FP32, capi.TF32])
@pytest.mark.parametrize("Cin,Cout,P", [(64, 128, 1000), (160, 64, 777), (512, 1536, 300), (8, 20, 130)])
def test_gemm_epilogue(Cin, Cout, P, mode):
    A = rnd(P, Cin).cuda()
    Wt = (rnd(Cout, Cin, seed=3) / Cin ** 0.5).cuda()
    bias, scale = rnd(Cout, seed=4).cuda(), rnd(Cout, seed=5).cuda()
    res = rnd(P, Cout, seed=6).cuda()
    out = torch.empty(P, Cout, device="cuda")
    gemm([(A, 0, Cin)], Cin, Wt, P, Cout, mode=mode, bias=bias, scale=scale, act=capi.ACT_GELU, res=res, out=out,
         ldo=Cout)
    want = F.gelu(A.cpu() @ Wt.cpu().t() + bias.cpu()) * scale.cpu() + res.cpu()
    tol = TOL if mode == capi.FP32 else TF32_REL * want.abs().max()
    assert (out.cpu() - want).abs().max() < tol


@pytest.mark.parametrize("mode", [capi.FP32, capi.TF32])
def test_gemm_segments_inplace_residual(mode):
    P, ch, nseg = 500, 64, 6
    bufs = [rnd(P, 3 * ch, seed=i).cuda() for i in range(nseg)]       # segment = middle column block
    Wt = (rnd(128, nseg * ch, seed=9) / 20).cuda()
    x = rnd(P, 128, seed=10).cuda()
    want = x.cpu() + torch.cat([b.cpu()[:, ch:2 * ch] for b in bufs], 1) @ Wt.cpu().t()
    gemm([(b, ch, 3 * ch) for b in bufs], ch, Wt, P, 128, mode=mode, res=x, out=x, ldo=128)
    tol = TOL if mode == capi.FP32 else TF32_REL * want.abs().max()
    assert (x.cpu() - want).abs().max() < tol


@pytest.mark.parametrize("mode", [capi.FP32, capi.TF32])
@pytest.mark.parametrize("store", [capi.STORE_UNSHUFFLE2, capi.STORE_SHUFFLE2, capi.STORE_PLAIN])
def test_conv3x3_im2col(store, mode):
    B, Cin, H, W = 2, 32, 12, 20
    Cout = 16 if store == capi.STORE_UNSHUFFLE2 else 64
    x = rnd(B, Cin, H, W)
    w = rnd(Cout, Cin, 3, 3, seed=1) / (9 * Cin) ** 0.5
    y = F.conv2d(x, w, padding=1)
    if store == capi.STORE_UNSHUFFLE2:
        want = F.pixel_unshuffle(y, 2)
    elif store == capi.STORE_SHUFFLE2:
        want = F.pixel_shuffle(y, 2)
    else:
        want = y
    Bo, Co, Ho, Wo = want.shape
    out = torch.empty(Bo, Ho, Wo, Co, device="cuda")
    wp = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous().cuda()
    gemm([(nhwc(x), 0, Cin)], Cin, wp, B * H * W, Cout, mode=mode, im2col=1, geom=(B, H, W), store=store, out=out,
         ldo=Co)
    tol = TOL if mode == capi.FP32 else TF32_REL * want.abs().max()
    assert (nchw(out) - want).abs().max() < tol


@pytest.mark.parametrize("fuse", [0, 1, 2])
def test_dwconv(fuse):
    B, C_, H, W = 2, 40, 9, 14
    x, w, b = rnd(B, C_, H, W), rnd(C_, 1, 3, 3, seed=1), rnd(C_, seed=2)
    y = F.conv2d(x, w, b, padding=1, groups=C_)
    if fuse == 1:
        y = F.gelu(y)
    elif fuse == 2:
        a, g = y.chunk(2, 1)
        y = F.gelu(a) * g
    Co = y.shape[1]
    out = torch.empty(B, H, W, Co, device="cuda")
    w9 = w.reshape(C_, 9).t().contiguous().cuda()
    call("turtle_dwconv3x3", dp(nhwc(x)), C_, w9.data_ptr(), dp(b), out.data_ptr(), Co, B, H, W,
         C_, fuse, 0, 1, 0, stream())
    assert (nchw(out) - y).abs().max() < TOL


def test_dwconv_patch_layout():
    B, C_, H, W, ws = 1, 8, 16, 24, 4
    x, w = rnd(B, C_, H, W), rnd(C_, 1, 3, 3, seed=1)
    want = orc.to_dilated_patches(F.conv2d(x, w, padding=1, groups=C_), ws)
    out = torch.empty(B, (H // ws) * (W // ws), ws * ws * C_, device="cuda")
    w9 = w.reshape(C_, 9).t().contiguous().cuda()
    call("turtle_dwconv3x3", dp(nhwc(x)), C_, w9.data_ptr(), None, out.data_ptr(), C_, B, H, W, C_, 0, 1, ws,
         0, stream())
    assert (out.cpu() - want).abs().max() < TOL


@pytest.mark.parametrize("up", [1, 4])
def test_pack_frame_first_last(up):
    B, Hs, Ws = 2, 10, 13
    x = torch.rand(B, 2, 3, Hs, Ws)
    img = x[:, 1]
    if up == 4:
        img = F.interpolate(img, scale_factor=4, mode="bilinear")
    H, W = img.shape[-2:]
    Hp, Wp = H + (-H) % 32, W + (-W) % 32
    want = F.pad(img, (0, Wp - W, 0, Hp - H))
    xd = x.cuda()
    dst = torch.empty(B, Hp, Wp, 3, device="cuda")
    call("turtle_pack_frame", xd.data_ptr() + 4 * 3 * Hs * Ws, 2 * 3 * Hs * Ws, dst.data_ptr(), B, 3, Hs, Ws, Hp, Wp,
         up, stream())
    assert (nchw(dst) - want).abs().max() < 1e-6
    # first conv
    w1 = rnd(16, 3, 3, 3, seed=1)
    y1 = torch.empty(B, Hp, Wp, 16, device="cuda")
    call("turtle_conv3x3_first", dst.data_ptr(), dp(w1), None, y1.data_ptr(), B, Hp, Wp, 3, 16, stream())
    assert (nchw(y1) - F.conv2d(want, w1, padding=1)).abs().max() < TOL
    # last conv + bias + current + crop
    w2, b2 = rnd(3, 16, 3, 3, seed=2), rnd(3, seed=3)
    out = torch.empty(B, 3, H, W, device="cuda")
    call("turtle_conv3x3_last", y1.data_ptr(), dp(w2), dp(b2), dst.data_ptr(), 3, 0,
         out.data_ptr(), B, Hp, Wp, 16, 3, H, W, stream())
    ref = (F.conv2d(F.conv2d(want, w1, padding=1), w2, b2, padding=1) + want)[:, :, :H, :W]
    assert (out.cpu() - ref).abs().max() < 1e-4


@pytest.mark.parametrize("mode", [capi.FP32, capi.TF32])
@pytest.mark.parametrize("heads,ch,S", [(4, 64, 1), (2, 64, 3), (2, 16, 2), (1, 64, 2), (1, 64, 5), (2, 64, 4), (1, 64, 8)])
def test_channel_attention_chain(heads, ch, S, mode):
    """gram -> softmax -> fold -> apply  ==  softmax(q^ k^T * t) @ v -> project_out, with S key segments."""
    c, P = heads * ch, 700
    qkv = [rnd(P, 3 * c, seed=s).cuda() for s in range(S)]      # segment s: its own k,v; q from the last
    temp = (torch.rand(heads) + 0.5)
    Wo = rnd(c, c, seed=7) / c ** 0.5
    x = rnd(P, c, seed=8)
    # reference on CPU, NCHW-like [heads, ch, P]
    def rows(t):
        return t.cpu().t().reshape(heads, ch, P)
    q = orc.l2norm_rows(rows(qkv[-1][:, :c]))
    k = torch.cat([orc.l2norm_rows(rows(t[:, c:2 * c])) for t in qkv], 1)
    v = torch.cat([rows(t[:, 2 * c:]) for t in qkv], 1)
    attn = torch.softmax(q @ k.transpose(-1, -2) * temp.view(-1, 1, 1), -1)
    want = x + ((attn @ v).reshape(c, P).t() @ Wo.t())
    nsplit = 5 if S < 4 else 40          # (40 splits: the softmax kernel's 8-deep unrolled partial sums run too)
    g = torch.zeros(S, nsplit, heads, ch, ch, device="cuda")
    sqq = torch.zeros(S, nsplit, c, device="cuda")
    sqk = torch.zeros(S, nsplit, c, device="cuda")
    qd = qkv[-1]
    for s in range(S):
        call("turtle_chan_gram", qd.data_ptr(), 3 * c, ch, qkv[s].data_ptr() + 4 * c, 3 * c, ch, P, heads, ch, nsplit,
             g[s].data_ptr(), sqq[s].data_ptr(), sqk[s].data_ptr(), mode, stream())
    flags = torch.zeros(S, dtype=torch.int32, device="cuda")
    Pm = torch.empty(heads, ch, S * ch, device="cuda")
    inv = torch.empty(S, c, device="cuda")
    call("turtle_chan_softmax", g.data_ptr(), sqq.data_ptr(), sqk.data_ptr(), flags.data_ptr(), dp(temp),
         S, nsplit, heads, ch, Pm.data_ptr(), inv.data_ptr(), stream())
    assert (Pm.cpu().reshape(heads, ch, S * ch) - attn).abs().max() < (1e-5 if mode == capi.FP32 else 2e-4)
    M = torch.empty(c, S * c, device="cuda")
    call("turtle_chan_fold", Pm.data_ptr(), dp(Wo), S, heads, ch, M.data_ptr(), 0, stream())
    xd = x.cuda()
    segs = [(qkv[s], 2 * c + h * ch, 3 * c) for s in range(S) for h in range(heads)]
    gemm(segs, ch, M, P, c, res=xd, out=xd, ldo=c)
    assert (xd.cpu() - want).abs().max() < (5e-5 if mode == capi.FP32 else 1e-3)


@pytest.mark.parametrize("mode", [capi.FP32, capi.TF32])
@pytest.mark.parametrize("heads,ch,S,B", [(4, 64, 2, 3), (2, 16, 1, 5), (1, 64, 3, 2)])
def test_channel_attention_batched_equals_per_element(heads, ch, S, B, mode):
    """turtle_chan_{gram,softmax,fold}_b over B batch elements in one launch each == the per-element entry points."""
    c, P, nsplit = heads * ch, 520, 3
    maps = [rnd(B, P, 3 * c, seed=10 + s).cuda() for s in range(S)]          # segment s: [B, P, 3c]; q from the last
    temp, Wo = (torch.rand(heads) + 0.5).cuda(), (rnd(c, c, seed=7) / c ** 0.5).cuda()
    flags = torch.zeros(S, dtype=torch.int32, device="cuda")
    g = torch.zeros(B, S, nsplit, heads, ch, ch, device="cuda")
    sqq, sqk = torch.zeros(B, S, nsplit, c, device="cuda"), torch.zeros(B, S, nsplit, c, device="cuda")
    qd = maps[-1]
    for s in range(S):
        call("turtle_chan_gram_b", qd.data_ptr(), 3 * c, ch, P * 3 * c, maps[s].data_ptr() + 4 * c, 3 * c, ch, P * 3 * c, P, heads,
             ch, nsplit, g[0, s].data_ptr(), sqq[0, s].data_ptr(), sqk[0, s].data_ptr(), g[0].numel(), sqq[0].numel(), B, mode,
             stream())
    Pm, inv = torch.empty(B, heads, ch, S * ch, device="cuda"), torch.empty(B, S, c, device="cuda")
    call("turtle_chan_softmax_b", g.data_ptr(), sqq.data_ptr(), sqk.data_ptr(), flags.data_ptr(), temp.data_ptr(), S, nsplit,
         heads, ch, Pm.data_ptr(), inv.data_ptr(), g[0].numel(), sqq[0].numel(), B, stream())
    M = torch.empty(B, c, S * c, device="cuda")
    call("turtle_chan_fold_b", Pm.data_ptr(), Wo.data_ptr(), S, heads, ch, M.data_ptr(), 0, B, stream())
    for b in range(B):
        g1 = torch.zeros(S, nsplit, heads, ch, ch, device="cuda")
        q1, k1 = torch.zeros(S, nsplit, c, device="cuda"), torch.zeros(S, nsplit, c, device="cuda")
        for s in range(S):
            call("turtle_chan_gram", qd[b].data_ptr(), 3 * c, ch, maps[s][b].data_ptr() + 4 * c, 3 * c, ch, P, heads, ch, nsplit,
                 g1[s].data_ptr(), q1[s].data_ptr(), k1[s].data_ptr(), mode, stream())
        P1, i1 = torch.empty(heads, ch, S * ch, device="cuda"), torch.empty(S, c, device="cuda")
        call("turtle_chan_softmax", g1.data_ptr(), q1.data_ptr(), k1.data_ptr(), flags.data_ptr(), temp.data_ptr(), S, nsplit,
             heads, ch, P1.data_ptr(), i1.data_ptr(), stream())
        M1 = torch.empty(c, S * c, device="cuda")
        call("turtle_chan_fold", P1.data_ptr(), Wo.data_ptr(), S, heads, ch, M1.data_ptr(), 0, stream())
        assert torch.equal(g[b], g1) and torch.equal(sqq[b], q1) and torch.equal(sqk[b], k1)
        assert torch.equal(Pm[b], P1) and torch.equal(inv[b], i1) and torch.equal(M[b], M1)


@pytest.mark.parametrize("bias", [False, True])
@pytest.mark.parametrize("ws,D", [(4, 64), (8, 32), (16, 16), (4, 20), (8, 128), (4, 256), (4, 512)])
def test_sab_window_reduce(ws, D, bias):
    B, H, W = 2, 32, 48
    t, w = rnd(B, D, H, W), rnd(D, 1, ws, ws, seed=1)
    b = rnd(D, seed=2) if bias else None
    want = orc.l2norm_rows(F.conv2d(t, w, b, stride=ws, padding=1, groups=D).flatten(2).transpose(1, 2))
    N = (H // ws) * (W // ws)
    out = torch.empty(B, N, D, device="cuda")
    wk = w.reshape(D, -1).t().contiguous().cuda()
    call("turtle_sab_window_reduce", dp(nhwc(t)), D, wk.data_ptr(), dp(b) if bias else None, out.data_ptr(), N * D, B, H, W,
         D, ws, stream())
    assert (out.cpu() - want).abs().max() < TOL


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("Hg,Wg,D,F_", [(6, 8, 32, 2), (16, 16, 128, 4), (30, 54, 64, 3), (46, 80, 128, 2)])
def test_sab_select_and_aggregate(Hg, Wg, D, F_, tc):
    N = Hg * Wg
    q = orc.l2norm_rows(rnd(N, D))
    k = orc.l2norm_rows(rnd(F_, N, D, seed=1))
    tau = 0.83
    Wt, top, S = orc.sab_select_sparse(q, k, tau, Hg, Wg)
    idx = torch.empty(F_, N, capi.SAB_SLOTS, dtype=torch.int32, device="cuda")
    wgt = torch.empty(F_, N, capi.SAB_SLOTS, device="cuda")
    if tc:     # tcgen05 3xTF32 correlation, top-5 per TMEM lane
        nbytes = capi.load().turtle_sab_select_tc_workspace(F_, N, D)
        wsp = torch.empty(nbytes // 4 + 1, device="cuda")
        call("turtle_sab_select_tc", dp(q), dp(k), N * D, F_, Hg, Wg, D, dp(torch.tensor([tau])), 0, idx.data_ptr(),
             wgt.data_ptr(), wsp.data_ptr(), stream())
    else:
        call("turtle_sab_select", dp(q), dp(k), N * D, F_, Hg, Wg, D,
             dp(torch.tensor([tau])), 0, idx.data_ptr(), wgt.data_ptr(), 0, stream())
    idx_c, wgt_c = idx.cpu().long(), wgt.cpu()
    # top-5 sets identical (fp32 FMA order may differ from the CPU matmul only on sub-ulp near-ties)
    same = (idx_c[..., :5].sort(-1).values == top.sort(-1).values).all(-1)
    bad = (~same).nonzero()
    for f, i in bad.tolist():
        srt = S[f, i].sort(descending=True).values
        assert (srt[4] - srt[5]).abs() < 1e-6, f"genuine top-k mismatch at frame {f} row {i}"
    assert same.float().mean() > 0.999
    dense = torch.zeros(F_, N, N)
    live = idx_c >= 0
    dense.scatter_add_(-1, idx_c.clamp_min(0), wgt_c * live)
    assert (dense[same] - Wt[same]).abs().max() < 2e-6
    # aggregation
    ws, c = 2, 8
    V = rnd(F_, N, ws * ws * c, seed=2)
    y = torch.empty(F_, Hg * ws, Wg * ws, c, device="cuda")
    call("turtle_sab_aggregate", idx.data_ptr(), wgt.data_ptr(), dp(V), N * ws * ws * c, y.data_ptr(), F_,
         Hg, Wg, ws, c, 0, 0, stream())
    want = orc.from_dilated_patches(dense @ V, ws, c, Hg * ws, Wg * ws)      # [F,c,H,W]
    assert (y.cpu().permute(0, 3, 1, 2) - want).abs().max() < 1e-5


@pytest.mark.parametrize("h16", [False, True])
@pytest.mark.parametrize("rmode", [0, 1, 2])
@pytest.mark.parametrize("Hg,Wg,ws,c,F_", [(13, 37, 2, 64, 2), (46, 80, 4, 32, 3), (8, 16, 4, 64, 1), (20, 9, 2, 64, 2),
                                           (5, 3, 2, 128, 1), (7, 33, 2, 64, 1)])
def test_sab_aggregate_tc(Hg, Wg, ws, c, F_, rmode, h16):
    """Tensor-core aggregation (dense 16x24 key box per 8x16 query tile + far top-k gather) == the CUDA-core kernel."""
    N, D, Dv = Hg * Wg, 32, ws * ws * c
    q = orc.l2norm_rows(rnd(N, D))
    k = orc.l2norm_rows(rnd(F_, N, D, seed=1))
    idx = torch.empty(F_, N, capi.SAB_SLOTS, dtype=torch.int32, device="cuda")
    wgt = torch.empty(F_, N, capi.SAB_SLOTS, device="cuda")
    call("turtle_sab_select", dp(q), dp(k), N * D, F_, Hg, Wg, D, dp(torch.tensor([0.83])), 0, idx.data_ptr(), wgt.data_ptr(),
         0, stream())
    V = rnd(F_, N, Dv, seed=2)
    if h16:         # the fp16 copy of the rows the engine keeps next to the ring
        V = V.half().float().cuda()
    else:           # TF32-rounded rows, as the engine writes them
        V = ((V.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).cuda()
    want = torch.empty(F_, Hg * ws, Wg * ws, c, device="cuda")
    call("turtle_sab_aggregate", idx.data_ptr(), wgt.data_ptr(), V.data_ptr(), N * Dv, want.data_ptr(), F_, Hg, Wg, ws, c, 0, 0,
         stream())
    wsp = torch.empty(capi.load().turtle_sab_aggregate_tc_workspace(F_, Hg, Wg) // 4, device="cuda")
    y = torch.full((F_, Hg * ws, Wg * ws, c), float("nan"), device="cuda", dtype=torch.float16 if rmode == 2 else torch.float32)
    Vin = V.half() if h16 else V
    call("turtle_sab_aggregate_tc", idx.data_ptr(), wgt.data_ptr(), Vin.data_ptr(), int(h16), N * Dv, y.data_ptr(), F_, Hg, Wg, ws,
         c, rmode, wsp.data_ptr(), stream())
    err = (y.float() - want).abs().max().item()
    assert err < (4e-3 if rmode == 2 else 2e-3), err
    # the far keys (top-k outside the tile's box) must be part of the sum: without them the error is O(0.1)
    assert (y.float() - want).abs().mean().item() < 2e-4


@pytest.mark.parametrize("bias", [False, True])
def test_dwconv_patch_rows_fp16_copy(bias):
    """turtle_dwconv3x3_patch_rows == turtle_dwconv3x3(layout 1) bit for bit, plus the fp16 copy of the same rows."""
    NB, H, W, c, ws = 2, 24, 40, 64, 4
    x = rnd(NB, H, W, c).cuda()
    w9 = rnd(9, c, seed=1).cuda()
    b = rnd(c, seed=2).cuda() if bias else None
    N, Dv = (H // ws) * (W // ws), ws * ws * c
    want = torch.empty(NB, N, Dv, device="cuda")
    call("turtle_dwconv3x3", x.data_ptr(), c, w9.data_ptr(), b.data_ptr() if bias else None, want.data_ptr(), c, NB, H, W, c, 0, 1,
         ws, 0, stream())
    y = torch.full((NB, N, Dv), float("nan"), device="cuda")
    y16 = torch.full((NB, N, Dv), float("nan"), device="cuda", dtype=torch.float16)
    call("turtle_dwconv3x3_patch_rows", x.data_ptr(), c, w9.data_ptr(), b.data_ptr() if bias else None, y.data_ptr(),
         y16.data_ptr(), NB, H, W, c, ws, stream())
    assert torch.equal(y, want)
    assert torch.equal(y16, want.half())


def test_library_is_loaded_from_tree():
    lib = capi.load()
    assert lib.turtle_abi_version() >= 1
    with open("/proc/self/maps") as f:
        assert "libturtle_b200.so" in f.read()
