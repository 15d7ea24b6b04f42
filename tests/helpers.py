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
