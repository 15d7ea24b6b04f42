"""Generate tests/golden/*.npz from the REFERENCE implementation (TEST INFRASTRUCTURE).

Runs only where /root/reference exists (the build container).  It imports the reference arch
files *by path* (they need only torch + einops), runs them on seeded synthetic
clips, checks oracle/turtle_oracle.py and the package's parameter tree against them, and
writes small fixtures that travel with the repo:

  full_<variant>_<gates>.npz   full-size yml config, weights from ``torch.manual_seed(10)`` default
                               init (+ optional live-gate randomisation), 3 frames of 96x128;
                               holds the input clip, the reference outputs, top-5 indices of every
                               SAB call, cache checksums, and a checksum of the weights so a torch
                               version whose init RNG differs is detected instead of mis-compared.
  tiny_<variant>.npz           dim-8 reduced config with its weights stored verbatim.

Usage:  python oracle/make_golden.py            (about a minute on 8 cores)
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("TURTLE_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")

from oracle.turtle_oracle import ArchSpec, Oracle, randomize_gates  # noqa: E402

ARCH_FILES = {"t1": "turtle_t1_arch", "t0": "turtle_arch", "super": "turtlesuper_t1_arch"}
YML = {"t1": "Turtle_Deblur_Gopro.yml", "t0": "Turtle_Derain.yml", "super": "Turtle_SR_MVSR.yml"}


def load_ref_module(variant):
    name = ARCH_FILES[variant]
    spec = importlib.util.spec_from_file_location(
        "ref_" + name, os.path.join(REF, "basicsr/models/archs", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_opt(variant):
    with open(os.path.join(REF, "options", YML[variant])) as f:
        return yaml.safe_load(f)


def tiny_opt(opt):
    o = dict(opt)
    o.update(dim=8, Enc_blocks=[1, 1, 2], Middle_blocks=2, Dec_blocks=[2, 1, 1], num_refinement_blocks=1,
             num_heads=[1, 1, 2, 4])
    return o


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


gaps = []


def run_reference(net, clip, spy_topk):
    """VRM:110-129 loop on the reference module, recording torch.topk indices (SAB only) and, in ``gaps``, the
    difference between the 5th and 6th largest score of each row."""
    gaps.clear()
    tops = []
    orig = torch.topk

    def spy(inp, k, dim=-1, **kw):
        r = orig(inp, k, dim=dim, **kw)
        tops.append(r.indices.clone())
        six = orig(inp, min(k + 1, inp.shape[dim]), dim=dim).values       # 5th/6th score gap: classifies mismatches
        gaps.append((six[..., k - 1] - six[..., k]).float() if six.shape[-1] > k else torch.zeros_like(six[..., 0]))
        return r

    outs, k, v = [], None, None
    caches = []
    torch.topk = spy if spy_topk else orig
    try:
        with torch.no_grad():
            for j in range(clip.shape[1]):
                pre = clip[:, j if j == 0 else j - 1]
                o, k, v = net(torch.stack([pre, clip[:, j]], 1).float(), k, v)
                outs.append(o)
                caches.append(([None if t is None else t.clone() for t in k],
                               [None if t is None else t.clone() for t in v]))
    finally:
        torch.topk = orig
    return torch.stack(outs, 1), caches, tops


def cache_digest(caches):
    """per frame, per slot: (sum, abs-sum) -- cheap fingerprints of the returned caches."""
    rows = []
    for ks, vs in caches:
        for t in list(ks) + list(vs):
            rows.append([0.0, 0.0] if t is None else [float(t.double().sum()), float(t.double().abs().sum())])
    return np.asarray(rows, dtype=np.float64)


def make_case(variant, gates, tiny, frames, H, W, seed, overrides=None, tag=""):
    opt = load_opt(variant)
    if tiny:
        opt = tiny_opt(opt)
    if overrides:
        opt.update(overrides)
    mod = load_ref_module(variant)
    torch.manual_seed(10)                         # yml manual_seed: 10
    ref = mod.make_model(opt).eval()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    init_sum = sd_checksum(sd)
    if gates == "live":
        sd = randomize_gates(sd, seed=1234)
        ref.load_state_dict(sd, strict=True)

    # the package's parameter tree must draw the identical default init and accept the state dict
    from turtlevsr_b200.archs import create_video_model
    torch.manual_seed(10)
    mine = create_video_model(opt)
    mine_sd = mine.state_dict()
    assert list(mine_sd.keys()) == list(ref.state_dict().keys()), "state-dict key order/name mismatch"
    assert sd_checksum({k: v for k, v in mine_sd.items()}) == init_sum, "default init differs from reference"
    mine.load_state_dict(sd, strict=True)

    g = torch.Generator().manual_seed(seed)
    if variant == "super":
        clip = torch.rand(1, frames, 3, H // 4, W // 4, generator=g)
    else:
        clip = torch.rand(1, frames, 3, H, W, generator=g)

    ref_out, ref_caches, ref_tops = run_reference(ref, clip, spy_topk=True)

    orc = Oracle(ArchSpec.from_opt(opt), sd)
    orc.trace = {}
    o_out, ok, ov = orc.run_clip(clip)
    err = (o_out - ref_out).abs().max().item()
    # caches of the last frame
    cerr = 0.0
    for a, b in zip(list(ok) + list(ov), list(ref_caches[-1][0]) + list(ref_caches[-1][1])):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.shape == b.shape, (a.shape, b.shape)
            cerr = max(cerr, (a - b).abs().max().item())
    o_tops = [r["topk"] for key in orc.trace for r in orc.trace[key]]
    # oracle trace is grouped per SAB module, reference spy is in call order: compare as sorted sets per call
    ref_sorted = sorted([t.sort(-1).values.flatten().tolist() for t in ref_tops])
    orc_sorted = sorted([t.sort(-1).values.flatten().tolist() for t in o_tops])
    topk_equal = ref_sorted == orc_sorted
    print(f"[{variant} {'tiny' if tiny else 'full'} {gates}] oracle vs reference: out max|d|={err:.3e} "
          f"cache max|d|={cerr:.3e} topk identical={topk_equal}  out range [{ref_out.min():.3f},{ref_out.max():.3f}]")
    assert err < 2e-5 and cerr < 2e-5, "oracle does not reproduce the reference"

    # top-k per SAB module in forward order (dec3, dec2, dec1) per frame, from the reference spy
    n_sab = 3
    tops = {}
    for i, t in enumerate(ref_tops):
        fr, lvl = divmod(i, n_sab)
        tops[f"topk_f{fr}_l{lvl}"] = t[0, :, 0].to(torch.int32).numpy()       # [F,N,5]
    ref_gaps = list(gaps)                           # the oracle below calls torch.topk un-spied
    for i, gp in enumerate(ref_gaps):
        fr, lvl = divmod(i, n_sab)
        tops[f"gap_f{fr}_l{lvl}"] = gp[0, :, 0].numpy()                       # [F,N]
    data = dict(
        clip=clip.numpy(), ref_out=ref_out.numpy(), cache_digest=cache_digest(ref_caches),
        init_checksum=np.array(init_sum), gates=np.array(gates), variant=np.array(variant),
        opt_yaml=np.array(yaml.safe_dump({k: v for k, v in opt.items()
                                          if not isinstance(v, dict)})),
        oracle_err=np.array(err), **tops)
    if tiny:
        for k, v in sd.items():
            data["w::" + k] = v.numpy()
    name = f"{'tiny' if tiny else 'full'}_{variant}_{gates}{tag}.npz"
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name), **data)
    print("   wrote", name, os.path.getsize(os.path.join(OUT, name)) // 1024, "KiB")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    make_case("t1", "init", tiny=False, frames=3, H=96, W=128, seed=7)
    make_case("t1", "live", tiny=False, frames=4, H=96, W=128, seed=8)
    make_case("t1", "live", tiny=True, frames=5, H=64, W=96, seed=9)
    make_case("super", "live", tiny=True, frames=3, H=128, W=128, seed=11)
    make_case("t0", "live", tiny=True, frames=3, H=64, W=64, seed=12)
    # the option switches no shipped yml uses: BiasFree LayerNorm (T1:67-81), both frames as input (T1:1059-1061),
    # conv biases (opt['bias'])
    make_case("t1", "live", tiny=True, frames=4, H=64, W=64, seed=13, tag="_biasfree_bothinputs",
              overrides=dict(LayerNorm_type="BiasFree", use_both_input=True))
    make_case("t1", "live", tiny=True, frames=4, H=64, W=64, seed=14, tag="_convbias", overrides=dict(bias=True))
    make_case("t0", "live", tiny=True, frames=3, H=64, W=64, seed=15, tag="_convbias", overrides=dict(bias=True))
