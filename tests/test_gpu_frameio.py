"""GPU parity of the frame-side kernels (csrc/frameio.cu) against the oracle restatements of the reference's own
host code: 8-bit conversions (utils/img_util.py:42-102, INFN:262-276), PSNR / SSIM (INF:33-61, metrics/psnr_ssim.py),
tile gather / overlap-average (INF:172-246).  Integer work (uint8 images, tile copies) must be bit-exact; PSNR within
1e-6 dB relative; SSIM within 2e-6 on [0,1]-scaled data and 1e-5 on the 0..255-scaled "basicsr" flavour, whose fp32
variance terms E[x^2] - mu^2 cancel at magnitude 6.5e4 (the reference itself computes them in fp32; bars written here)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import metrics_oracle as mo  # noqa: E402
from oracle.make_golden_metrics import CASES, frame_pair  # noqa: E402
from turtlevsr_b200 import frameio  # noqa: E402


@pytest.mark.parametrize("H,W,C", [(48, 64, 3), (37, 53, 3), (720, 1280, 3), (16, 20, 1), (9, 7, 4)])
def test_u8_frame_roundtrip_and_quantisation(H, W, C):
    g = torch.Generator().manual_seed(H * W)
    img = torch.randint(0, 256, (H, W, C), dtype=torch.uint8, generator=g)
    f = frameio.u8_to_frame(img.cuda())
    want = img.permute(2, 0, 1).float() / 255
    assert torch.equal(f.cpu(), want)              # IEEE division by 255, bit for bit
    fb = frameio.u8_to_frame(img.cuda(), swap_rb=True)
    if C >= 3:
        assert torch.equal(fb[0], f[2]) and torch.equal(fb[2], f[0]) and torch.equal(fb[1], f[1])
    # quantise: un-clamped restored frame, round-half-even (tensor2img) and truncation (INFN:268-269), bit-exact
    x = (torch.rand(C, H, W, generator=g) * 1.4 - 0.2)
    x[0, 0, 0] = 0.5 / 255.0 + 1.0 / 255.0        # an exact .5 case after scaling is representable only approximately
    got = frameio.frame_to_u8(x.cuda()).cpu().numpy()
    if C == 3:
        assert np.array_equal(got, mo.tensor2img_u8(x))
    want_r = (x.clamp(0, 1).numpy().transpose(1, 2, 0) * 255.0).round().astype(np.uint8)
    assert np.array_equal(got, want_r)
    got_t = frameio.frame_to_u8(x.cuda(), round_half_even=False).cpu().numpy()
    assert np.array_equal(got_t, (x.clamp(0, 1).permute(1, 2, 0).numpy() * 255).astype(np.uint8))
    if C >= 3:
        assert np.array_equal(frameio.frame_to_u8(x.cuda(), swap_rb=True).cpu().numpy()[..., :3], got[..., 2::-1])
    # u8 -> frame -> u8 is the identity
    assert torch.equal(frameio.frame_to_u8(f).cpu(), img)


@pytest.mark.parametrize("flavour", ["inference", "basicsr", "float"])
def test_frame_metrics_match_reference_formulas(flavour):
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "metrics_golden.npz"))
    cols = list(z["columns"])
    for row, (seed, H, W, noise) in zip(z["cases"], CASES):
        out, gt = frame_pair(seed, H, W, noise)
        psnr, ssim = frameio.frame_metrics(out.cuda(), gt.cuda(), flavour)
        want_p, want_s = mo.frame_metrics(out, gt, flavour)
        tol = 1e-5 if flavour == "basicsr" else 2e-6
        assert abs(psnr - want_p) <= 1e-6 * abs(want_p), (flavour, seed, psnr, want_p)
        assert abs(ssim - want_s) <= tol, (flavour, seed, ssim, want_s)
        # and against the values the reference's own functions produced (oracle/make_golden_metrics.py)
        if flavour == "inference":
            assert abs(psnr - row[cols.index("inf_psnr")]) <= 1e-6 * abs(psnr)
            assert abs(ssim - row[cols.index("inf_ssim")]) <= 2e-6
        else:
            assert abs(ssim - row[cols.index("bsr_ssim" if flavour == "basicsr" else "flt_ssim")]) <= tol
    # identical frames: PSNR = +inf, SSIM = 1
    p, s = frameio.frame_metrics(gt.cuda(), gt.cuda(), flavour)
    assert p == float("inf") and abs(s - 1.0) < 1e-6


def test_frame_metrics_720p_and_accumulator():
    out, gt = frame_pair(7, 720, 1280, 0.03)
    acc = frameio.FrameMetrics("inference")
    for j in range(3):
        acc.add((out + 0.01 * j).cuda(), gt.cuda())
    r = acc.per_frame()
    want_p, want_s = mo.frame_metrics(out, gt, "inference")
    assert abs(r[0, 0] - want_p) <= 1e-6 * want_p and abs(r[0, 1] - want_s) <= 2e-6
    assert r[0, 3] == 3 * 720 * 1280 and r.shape == (3, 4)
    assert r[2, 0] < r[0, 0]                      # the disturbed frames score lower
    mp, ms = acc.mean()
    assert abs(mp - float(r[:, 0].mean())) < 1e-12


def test_tile_gather_and_blend_match_reference_loop():
    """INF:185-245 restated with torch ops (reflect pad, tile slices, E/W accumulation, clamp) vs the two kernels."""
    import ctypes as C
    import torch.nn.functional as F
    from turtlevsr_b200.capi import call
    g = torch.Generator().manual_seed(3)
    for (H, W, tile, ov) in [(90, 150, 64, 32), (720, 1280, 320, 128), (45, 77, 40, 8)]:
        prev, cur = torch.rand(1, 3, H, W, generator=g).cuda(), torch.rand(1, 3, H, W, generator=g).cuda()
        Hp, Wp = ((H + 8) // 8) * 8, ((W + 8) // 8) * 8
        padh, padw = (Hp - H if H % 8 else 0), (Wp - W if W % 8 else 0)
        pc, pp = F.pad(cur, (0, padw, 0, padh), "reflect"), F.pad(prev, (0, padw, 0, padh), "reflect")
        h, w = pc.shape[-2:]
        st = tile - ov
        ys, xs = list(range(0, h - tile, st)) + [h - tile], list(range(0, w - tile, st)) + [w - tile]
        want = torch.stack([torch.stack([pp[0, :, y:y + tile, x:x + tile], pc[0, :, y:y + tile, x:x + tile]])
                            for y in ys for x in xs])
        got = torch.empty_like(want)
        ya, xa = (C.c_int32 * len(ys))(*ys), (C.c_int32 * len(xs))(*xs)
        s = torch.cuda.current_stream().cuda_stream
        call("turtle_tile_gather", prev.data_ptr(), cur.data_ptr(), got.data_ptr(), 3, H, W, tile, ya, len(ys), xa,
             len(xs), s)
        assert torch.equal(got, want)
        tiles = torch.rand(len(ys) * len(xs), 3, tile, tile, generator=g).cuda() * 1.2 - 0.1
        E, Wt = torch.zeros(1, 3, h, w, device="cuda"), torch.zeros(1, 3, h, w, device="cuda")
        i = 0
        for y in ys:
            for x in xs:
                E[..., y:y + tile, x:x + tile].add_(tiles[i:i + 1])
                Wt[..., y:y + tile, x:x + tile].add_(1.0)
                i += 1
        want_b = torch.clamp(E.div_(Wt), 0, 1)
        got_b = torch.empty(1, 3, h, w, device="cuda")
        call("turtle_tile_blend", tiles.data_ptr(), got_b.data_ptr(), 3, h, w, tile, ya, len(ys), xa, len(xs), 1, s)
        assert (got_b - want_b).abs().max() <= 1.2e-7          # same sums in the same order; one division each


def test_streamed_u8_clip_matches_float_clip_quantised():
    """8-bit host clip in, 8-bit host clip out, on-device metrics: equals quantising the float pipeline's output."""
    from helpers import load_case
    from turtlevsr_b200.archs import create_video_model
    from turtlevsr_b200.clip import run_clip
    opt, sd, _, _, _ = load_case("tiny_t1_live.npz")
    net = create_video_model(opt)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval().set_precision("tf32")
    g = torch.Generator().manual_seed(11)
    clip8 = torch.randint(0, 256, (6, 32, 64, 3), dtype=torch.uint8, generator=g)
    gt8 = torch.randint(0, 256, (6, 32, 64, 3), dtype=torch.uint8, generator=g)
    clipf = (clip8.permute(0, 3, 1, 2).float() / 255)[None].cuda()
    want, _, _ = run_clip(net, clipf)
    acc = frameio.FrameMetrics("inference")
    got, k, v, last = frameio.run_clip_streamed_u8(net, clip8.pin_memory(), gt_u8=gt8.pin_memory(), metrics=acc)
    for j in range(6):
        assert np.array_equal(got[j].numpy(), mo.tensor2img_u8(want[0, j]))
    r = acc.per_frame()
    for j in (0, 5):
        wp, wssim = mo.frame_metrics(want[0, j].cpu(), gt8[j].permute(2, 0, 1).float() / 255, "inference")
        assert abs(r[j, 0] - wp) <= 1e-6 * wp and abs(r[j, 1] - wssim) <= 2e-6
