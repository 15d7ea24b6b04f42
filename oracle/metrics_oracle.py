"""CPU restatement of the reference's frame metrics and 8-bit conversions (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.  Each function cites the
reference code it follows; oracle/make_golden_metrics.py pins them against the reference's own functions
(executed from /root/reference) and commits the values under tests/golden/metrics_golden.npz.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from scipy.ndimage import gaussian_filter


def tensor2img_u8(t: torch.Tensor) -> np.ndarray:
    """utils/img_util.py:42-102 with rgb2bgr=False, out_type uint8: [C,H,W] float -> clamp(0,1) -> HWC -> (x*255).round()."""
    x = t.squeeze(0).float().detach().cpu().clamp(0, 1).numpy().transpose(1, 2, 0)
    return (x * 255.0).round().astype(np.uint8)


def psnr_u8(a: np.ndarray, b: np.ndarray) -> float:
    """inference.py:52-61 (calc_PSNR): uint8 images, float64 MSE, peak 255."""
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 20 * math.log10(255.0 / math.sqrt(mse))


def ssim_inference(a: np.ndarray, b: np.ndarray, sd: float = 1.5, C1: float = 0.01 ** 2, C2: float = 0.03 ** 2) -> float:
    """inference.py:33-50 (ssim_calculate): float32 images / 255, scipy gaussian_filter over ALL axes of the HWC array
    (rows, columns and the three channels), default mode 'reflect', truncate 4.0."""
    x = np.array(a, dtype=np.float32) / 255
    y = np.array(b, dtype=np.float32) / 255
    mu1, mu2 = gaussian_filter(x, sd), gaussian_filter(y, sd)
    s1 = gaussian_filter(x * x, sd) - mu1 * mu1
    s2 = gaussian_filter(y * y, sd) - mu2 * mu2
    s12 = gaussian_filter(x * y, sd) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return float(np.mean(m))


def psnr_basicsr(a: np.ndarray, b: np.ndarray, peak=None) -> float:
    """metrics/psnr_ssim.py:13-68 (calculate_psnr, crop_border 0, no Y channel): float64 MSE,
    peak 1 if img1.max() <= 1 else 255 (psnr_ssim.py:66) unless ``peak`` fixes it."""
    a, b = a.astype(np.float64), b.astype(np.float64)
    mse = np.mean((a - b) ** 2)
    if mse == 0:
        return float("inf")
    if peak is None:
        peak = 1.0 if a.max() <= 1 else 255.0
    return float(20.0 * np.log10(peak / np.sqrt(mse)))


def _gauss11() -> np.ndarray:
    """cv2.getGaussianKernel(11, 1.5) (psnr_ssim.py:137): normalised exp(-x^2 / (2 sigma^2)) for sigma > 0."""
    k = np.exp(-0.5 * (np.arange(-5, 6, dtype=np.float64) ** 2) / 1.5 ** 2)
    return k / k.sum()


def ssim_basicsr_3d(a: np.ndarray, b: np.ndarray, max_value: float) -> float:
    """metrics/psnr_ssim.py:136-180 (_generate_3d_gaussian_kernel + _ssim_3d): the HWC image is one 3-D volume filtered
    by an 11x11x11 Gaussian (float32 Conv3d, padding 5, padding_mode 'replicate'); SSIM map averaged over the volume."""
    C1, C2 = (0.01 * max_value) ** 2, (0.03 * max_value) ** 2
    g = _gauss11()
    w3 = torch.tensor(np.stack([np.outer(g, g) * k for k in g], 0)).float()[None, None]

    def filt(v):
        v = torch.nn.functional.pad(v[None, None], (5, 5, 5, 5, 5, 5), mode="replicate")
        return torch.nn.functional.conv3d(v, w3)[0, 0]

    x = torch.tensor(a.astype(np.float64)).float()
    y = torch.tensor(b.astype(np.float64)).float()
    mu1, mu2 = filt(x), filt(y)
    s1 = filt(x * x) - mu1 * mu1
    s2 = filt(y * y) - mu2 * mu2
    s12 = filt(x * y) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    return float(m.mean())


def frame_metrics(restored: torch.Tensor, gt: torch.Tensor, flavour: str):
    """-> (psnr, ssim) as the reference computes them for one frame pair ([C,H,W] float tensors).
    'inference': INF:313-327; 'basicsr': VRM:171-200 (tensor2img then calculate_psnr / calculate_ssim);
    'float': the psnr_ssim.py formulas on un-quantised data of nominal range [0,1] with max_value FIXED to 1 (the
    reference switches to 255 as soon as one restored value exceeds 1.0, psnr_ssim.py:66 -- a data-dependent rule that
    the un-clamped network output would trip; the trainer never calls it on such data, it quantises first = 'basicsr')."""
    if flavour == "inference":
        a, b = tensor2img_u8(restored), tensor2img_u8(gt)
        return psnr_u8(a, b), ssim_inference(a, b)
    if flavour == "basicsr":
        a, b = tensor2img_u8(restored), tensor2img_u8(gt)
        return psnr_basicsr(a, b), ssim_basicsr_3d(a, b, 255.0)
    a = restored.detach().cpu().numpy().transpose(1, 2, 0)
    b = gt.detach().cpu().numpy().transpose(1, 2, 0)
    return psnr_basicsr(a, b, peak=1.0), ssim_basicsr_3d(a, b, 1.0)
