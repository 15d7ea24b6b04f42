"""Shared test helpers (CPU side): fixture loading, seeded model construction."""
import hashlib
import os

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def load_case(name):
    """-> (opt dict, state_dict, clip, ref_out, raw npz).  Weights are either stored ("w::" keys,
    tiny cases) or re-drawn from ``torch.manual_seed(10)`` default init and checked against the
    checksum recorded when the reference produced the fixture."""
    from oracle.turtle_oracle import randomize_gates
    from turtlevsr_b200.archs import create_video_model
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    opt = yaml.safe_load(str(z["opt_yaml"]))
    stored = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    if stored:
        sd = stored
    else:
        torch.manual_seed(10)
        net = create_video_model(opt)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        if sd_checksum(sd) != str(z["init_checksum"]):
            import pytest
            pytest.skip("this torch build draws a different default init than the fixture's")
        if str(z["gates"]) == "live":
            sd = randomize_gates(sd, seed=1234)
    return opt, sd, torch.from_numpy(z["clip"]), torch.from_numpy(z["ref_out"]), z


# ----------------------------------------------------------------------------------------
# named BASELINE.json configurations (fixtures of oracle/make_golden_named.py)
# ----------------------------------------------------------------------------------------
NEAR_TIE = 2e-6     # |5th - 6th| reference score below which a top-5 flip is fp32 summation-order noise (scores are
                    # tau * cosine, |S| <= 1.5; one fp32 ulp at 1.0 is 1.2e-7 and a 512-term dot product plus the conv
                    # chain in front of it reorders ~10 of them).  Anything above is a genuine error.


def load_named(name):
    """-> (opt, state_dict, clip, clean clip, npz).  The clip is re-drawn from the fixture's seed and checked
    against the sha256 recorded when the reference produced the fixture."""
    import pytest
    from oracle.turtle_oracle import randomize_gates
    from turtlevsr_b200.archs import create_video_model
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    opt = yaml.safe_load(str(z["opt_yaml"]))
    stored = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    if stored:
        sd = stored
    else:
        torch.manual_seed(10)
        net = create_video_model(opt)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        if sd_checksum(sd) != str(z["init_checksum"]):
            pytest.skip("this torch build draws a different default init than the fixture's")
        sd = randomize_gates(sd, seed=1234)
    frames, H, W = int(z["frames"]), int(z["H"]), int(z["W"])
    lr = 4 if str(z["variant"]) == "super" else 1
    g = torch.Generator().manual_seed(int(z["seed"]))
    clean = torch.rand(1, frames, 3, H // lr, W // lr, generator=g)
    sigma = float(z["noise_sigma"])
    clip = clean + torch.randn(clean.shape, generator=g) * sigma if sigma else clean
    if hashlib.sha256(clip.contiguous().numpy().tobytes()).hexdigest() != str(z["clip_sha"]):
        pytest.skip("this torch build draws a different synthetic clip than the fixture's")
    return opt, sd, clip, clean, z


def frame_error(o, z, j):
    """max|out - reference| of frame j over everything the fixture kept of the reference's output:
    strided subsample, full-resolution crops, whole frames where stored; plus the largest deviation of a
    16x16 block MEAN (covers every pixel).  o: [C,H,W] CPU tensor."""
    st, cr = int(z["stride"]), int(z["crop"])
    err = float((o[:, ::st, ::st] - torch.from_numpy(z["sub"][j])).abs().max())
    for i, (y, x) in enumerate(z["crop_origins"].reshape(-1, 2)):
        err = max(err, float((o[:, y:y + cr, x:x + cr] - torch.from_numpy(z["crops"][j, i])).abs().max()))
    if f"full_{j}" in z.files:
        err = max(err, float((o - torch.from_numpy(z[f"full_{j}"])).abs().max()))
    C, H, W = o.shape
    Hb, Wb = H // 16, W // 16
    blk = o[:, :Hb * 16, :Wb * 16].double().reshape(C, Hb, 16, Wb, 16).sum(dim=(2, 4))
    berr = float((blk - torch.from_numpy(z["blocks"][j])).abs().max()) / 256.0
    return err, berr


def classify_topk(z, fr, lvl, got_idx, tag=""):
    """Compare the top-5 index SETS of one StateAlignBlock call with the reference's torch.topk.
    got_idx: [F,N,>=5] integer tensor.  Every mismatching row is classified by the reference's 5th/6th
    score gap: exact tie (gap == 0), near-tie (gap <= NEAR_TIE), genuine error.  -> (rows, ties, near, genuine)"""
    want = np.sort(z[f"topk_f{fr}_l{lvl}"].astype(np.int64), -1)
    got = np.sort(got_idx[..., :5].cpu().numpy().astype(np.int64), -1)
    assert want.shape == got.shape, (want.shape, got.shape)
    gap = z[f"gap_f{fr}_l{lvl}"] if f"gap_f{fr}_l{lvl}" in z.files else None
    bad = np.argwhere((want != got).any(-1))
    ties = near = genuine = 0
    for f, n in bad:
        gp = float(gap[f, n]) if gap is not None else float("nan")
        kind = "exact tie" if gp == 0 else ("near-tie" if abs(gp) <= NEAR_TIE else "GENUINE")
        ties += kind == "exact tie"
        near += kind == "near-tie"
        genuine += kind == "GENUINE"
        print(f"   top-5 mismatch {tag} frame {fr} level {lvl} f={f} row={n}: reference 5th-6th gap {gp:.3e} -> {kind}; "
              f"want {want[f, n].tolist()} got {got[f, n].tolist()}")
    return want.shape[0] * want.shape[1], ties, near, genuine
