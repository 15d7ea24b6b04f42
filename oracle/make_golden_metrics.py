"""Pin oracle/metrics_oracle.py against the REFERENCE's own metric functions (TEST INFRASTRUCTURE).

``basicsr/inference.py`` and ``basicsr/metrics/psnr_ssim.py`` cannot be imported here (matplotlib / skimage are
absent, and _ssim_3d insists on ``.cuda()``), so the functions themselves are lifted out of the reference's source
files with ``ast`` and executed -- unmodified -- in a namespace that provides numpy / scipy / cv2 / torch, with
``.cuda()`` turned into a no-op.  Their results on seeded frame pairs are written to tests/golden/metrics_golden.npz
together with the inputs' seeds; the oracle must reproduce them.

Usage:  python oracle/make_golden_metrics.py
"""
from __future__ import annotations

import ast
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("TURTLE_REFERENCE", "/root/reference")

from oracle import metrics_oracle as mo  # noqa: E402


def lift(path, names, ns):
    """exec the top-level function definitions ``names`` of a reference source file into ``ns``."""
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(body) == len(names), (path, names)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)


def frame_pair(seed, H, W, noise):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(3, H, W, generator=g)
    # smooth the clean frame a little so that SSIM is not degenerate, then disturb it (un-clamped, like a restored frame)
    gt = torch.nn.functional.avg_pool2d(gt[None], 3, 1, 1)[0]
    out = gt + torch.randn(3, H, W, generator=g) * noise
    return out, gt


CASES = [(1, 48, 64, 0.02), (2, 67, 45, 0.1), (3, 128, 160, 0.005), (4, 16, 16, 0.3)]

if __name__ == "__main__":
    import cv2
    from scipy.ndimage import gaussian_filter
    ns_inf = dict(np=np, math=math, gaussian_filter=gaussian_filter)
    lift(os.path.join(REF, "basicsr/inference.py"), ["ssim_calculate", "calc_PSNR"], ns_inf)
    ns_m = dict(np=np, cv2=cv2, torch=torch)
    lift(os.path.join(REF, "basicsr/metrics/psnr_ssim.py"),
         ["_3d_gaussian_calculator", "_generate_3d_gaussian_kernel", "_ssim_3d"], ns_m)
    ns_t = dict(np=np, cv2=cv2, torch=torch, math=math, make_grid=None)
    lift(os.path.join(REF, "basicsr/utils/img_util.py"), ["tensor2img"], ns_t)
    torch.Tensor.cuda = lambda self, *a, **k: self            # _ssim_3d moves everything to the GPU
    torch.nn.Module.cuda = lambda self, *a, **k: self

    rows = []
    for seed, H, W, noise in CASES:
        out, gt = frame_pair(seed, H, W, noise)
        # (the reference's tensor2img clamps its argument IN PLACE -- img_util.py:73 clamp_ on a view -- so it gets copies)
        a, b = ns_t["tensor2img"](out.clone(), rgb2bgr=False), ns_t["tensor2img"](gt.clone(), rgb2bgr=False)
        assert np.array_equal(a, mo.tensor2img_u8(out)) and np.array_equal(b, mo.tensor2img_u8(gt))
        ref = dict(
            inf_psnr=ns_inf["calc_PSNR"](a, b), inf_ssim=float(ns_inf["ssim_calculate"](a, b)),
            bsr_ssim=float(ns_m["_ssim_3d"](a.astype(np.float64), b.astype(np.float64), 255)),
            flt_ssim=float(ns_m["_ssim_3d"](out.numpy().transpose(1, 2, 0), gt.numpy().transpose(1, 2, 0), 1)))
        mine = dict(zip(("inf_psnr", "inf_ssim"), mo.frame_metrics(out, gt, "inference")))
        mine["bsr_ssim"] = mo.frame_metrics(out, gt, "basicsr")[1]
        mine["flt_ssim"] = mo.frame_metrics(out, gt, "float")[1]
        for k in ref:
            assert abs(ref[k] - mine[k]) <= 1e-6 * max(1.0, abs(ref[k])), (seed, k, ref[k], mine[k])
        print(f"seed {seed} {W}x{H}: " + "  ".join(f"{k}={v:.6f}" for k, v in ref.items()) + "   oracle agrees")
        rows.append([seed, H, W, noise, ref["inf_psnr"], ref["inf_ssim"], ref["bsr_ssim"], ref["flt_ssim"]])
    np.savez(os.path.join(ROOT, "tests", "golden", "metrics_golden.npz"), cases=np.array(rows, dtype=np.float64),
             columns=np.array(["seed", "H", "W", "noise", "inf_psnr", "inf_ssim", "bsr_ssim", "flt_ssim"]))
    print("wrote tests/golden/metrics_golden.npz")
