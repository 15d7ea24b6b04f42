"""Stage the reference's own arch files for the reference arm (TEST / BASELINE INFRASTRUCTURE).

The reference is pure Python: its hot path is three arch files that need only torch + einops.  This recipe copies
them -- unmodified -- together with the shipped option files from /root/reference into ``oracle/_ref/`` (git-ignored:
reference sources never enter the repository's history; not gpurun-ignored: the staged copy travels to the GPU box
like the built .so), so that ``bench.py --impl reference`` and the ``gpu_eager_baseline`` leg run the REFERENCE'S
code, not the oracle port.  Runs from ``__graft_entry__.build()`` whenever /root/reference is present.

    python oracle/build_ref.py
"""
from __future__ import annotations

import importlib.util
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("TURTLE_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
ARCHS = ["turtle_arch.py", "turtle_t1_arch.py", "turtlesuper_t1_arch.py"]
VARIANT_FILE = {"t0": "turtle_arch", "t1": "turtle_t1_arch", "super": "turtlesuper_t1_arch"}


def build() -> bool:
    """-> True if oracle/_ref is staged (now or earlier), False if there is no reference tree to stage from."""
    src_arch = os.path.join(REF, "basicsr", "models", "archs")
    if not os.path.isdir(src_arch):
        return os.path.isdir(os.path.join(DST, "archs"))
    os.makedirs(os.path.join(DST, "archs"), exist_ok=True)
    os.makedirs(os.path.join(DST, "options"), exist_ok=True)
    for f in ARCHS:
        shutil.copyfile(os.path.join(src_arch, f), os.path.join(DST, "archs", f))
    for f in sorted(os.listdir(os.path.join(REF, "options"))):
        if f.endswith(".yml"):
            shutil.copyfile(os.path.join(REF, "options", f), os.path.join(DST, "options", f))
    return True


def available() -> bool:
    return all(os.path.exists(os.path.join(DST, "archs", f)) for f in ARCHS)


def load_arch(variant: str):
    """The staged reference arch module ('t0' | 't1' | 'super'), loaded by path."""
    name = VARIANT_FILE[variant]
    spec = importlib.util.spec_from_file_location("turtle_ref_" + name, os.path.join(DST, "archs", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_opt(yml: str) -> dict:
    import yaml
    with open(os.path.join(DST, "options", yml)) as f:
        opt = yaml.safe_load(f)
    for k, v in list(opt.items()):            # the Davis yml's MEST / CTS build no model as shipped
        if v == "MEST":
            opt[k] = "CHM"
        elif v == "CTS":
            opt[k] = "FHR"
    return opt


if __name__ == "__main__":
    print("staged" if build() else "no reference tree", DST)
