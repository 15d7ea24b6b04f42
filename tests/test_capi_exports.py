"""CPU: the C-ABI library builds/loads here and exports every symbol include/turtle_b200.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

from turtlevsr_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "turtle_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(turtle_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = capi.load()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in turtle_b200.h but not exported by {capi.lib_path()}"


def test_binding_table_matches_header():
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_abi_version_and_struct_layout():
    lib = capi.load()
    assert lib.turtle_abi_version() >= 1
    assert b"sm_100a" in lib.turtle_build_info()
    # TurtleGemmArgs: 48 segment pointers + 48 pitches must match TURTLE_MAX_SEG in the header
    hdr = open(os.path.join(ROOT, "include", "turtle_b200.h")).read()
    assert int(re.search(r"#define TURTLE_MAX_SEG (\d+)", hdr).group(1)) == capi.MAX_SEG
    assert int(re.search(r"#define TURTLE_SAB_SLOTS (\d+)", hdr).group(1)) == capi.SAB_SLOTS
    assert lib.turtle_abi_version() == 3
    assert ctypes.sizeof(capi.GemmArgs) == lib.turtle_sizeof_gemm_args() == 728


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = capi.load()
    # NULL pointers / bad shapes return TURTLE_EINVAL before any launch is attempted
    assert lib.turtle_layernorm(None, 64, None, None, None, 64, 10, 64, 0, None) == -1
    assert lib.turtle_dwconv3x3(None, 64, None, None, None, 64, 1, 8, 8, 64, 0, 0, 1, 0, None) == -1
    a = capi.GemmArgs()
    assert lib.turtle_gemm(ctypes.byref(a), None) == -1
