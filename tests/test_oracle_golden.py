"""CPU: the oracle restatement reproduces the fixtures the REFERENCE produced (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from helpers import load_case
from oracle.turtle_oracle import (ArchSpec, Oracle, clipped_softmax_rows, from_dilated_patches,
                                  local_l1_mask, to_dilated_patches)

CASES = ["tiny_t1_live.npz", "tiny_super_live.npz", "tiny_t0_live.npz", "full_t1_init.npz"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    opt, sd, clip, ref_out, z = load_case(name)
    if name.startswith("full"):
        clip, ref_out = clip[:, :2], ref_out[:, :2]          # keep the CPU suite short
    orc = Oracle(ArchSpec.from_opt(opt), sd)
    orc.trace = {}
    out, ks, vs = orc.run_clip(clip)
    assert out.shape == ref_out.shape
    assert (out - ref_out).abs().max().item() < 2e-5
    # top-5 indices recorded from the reference's own torch.topk calls
    per_mod = list(orc.trace.values())               # dec3, dec2, dec1
    for lvl, recs in enumerate(per_mod):
        for fr, rec in enumerate(recs):
            want = np.sort(z[f"topk_f{fr}_l{lvl}"], axis=-1)
            got = np.sort(rec["topk"][0, :, 0].numpy(), axis=-1)
            assert (want == got).all()


def test_cache_protocol_shapes():
    opt, sd, clip, _, _ = load_case("tiny_t1_live.npz")
    orc = Oracle(ArchSpec.from_opt(opt), sd)
    _, ks, vs = orc.run_clip(clip)
    assert len(ks) == 8 and len(vs) == 8
    assert ks[0] is None and ks[1] is None and ks[2] is None
    H, W = clip.shape[-2:]
    N = (H // 16) * (W // 16)
    dim = opt["dim"]
    assert ks[3].shape == (1, 4, 3 * dim * 8 // 4, (H // 8) * (W // 8))
    assert ks[5].shape == (1, 3, 1, N, 2 * dim * 4) and vs[5].shape == (1, 3, 1, N, 16 * dim * 4)
    assert ks[7].shape == (1, 2, 1, N, 2 * dim) and vs[7].shape == (1, 2, 1, N, 256 * dim)


def test_dilated_patch_roundtrip():
    v = torch.randn(2, 6, 8, 12)
    p = to_dilated_patches(v, 4)
    assert p.shape == (2, 6, 96)
    # element (p1,p2,d) of patch (i,j) is v[d, p1*Hg+i, p2*Wg+j]   
    Hg, Wg = 2, 3
    i, j, p1, p2, d = 1, 2, 3, 1, 4
    assert p[1, i * Wg + j, (p1 * 4 + p2) * 6 + d] == v[1, d, p1 * Hg + i, p2 * Wg + j]
    assert torch.equal(from_dilated_patches(p, 4, 6, 8, 12), v)


def test_clipped_softmax_and_mask():
    z = torch.tensor([[0.0, 1.0, 2.0, 0.0]])
    w = clipped_softmax_rows(z)
    assert w[0, 0] == 0 and w[0, 3] == 0 and abs(w.sum().item() - 1) < 1e-6
    m = local_l1_mask(46, 80, 4)
    assert int(m.sum(1).max()) == 41 and int(m.sum(1).min()) == 15
