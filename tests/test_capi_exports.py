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
    assert lib.turtle_abi_version() == 5
    assert ctypes.sizeof(capi.GemmArgs) == lib.turtle_sizeof_gemm_args() == 752


def test_bad_arguments_are_rejected_without_a_gpu():
    lib = capi.load()
    # NULL pointers / bad shapes return TURTLE_EINVAL before any launch is attempted
    assert lib.turtle_layernorm(None, 64, None, None, None, 64, 10, 64, 0, None) == -1
    assert lib.turtle_dwconv3x3(None, 64, None, None, None, 64, 1, 8, 8, 64, 0, 0, 1, 0, None) == -1
    a = capi.GemmArgs()
    assert lib.turtle_gemm(ctypes.byref(a), None) == -1


def test_every_kernel_entry_point_has_a_torch_custom_op():
    """ops.py: one torch.ops.turtle_b200.<name> per kernel-launching prototype of the header, schema derived from it."""
    import torch
    from turtlevsr_b200 import ops
    hdr = open(os.path.join(ROOT, "include", "turtle_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    launching = sorted(set(m.group(1) for m in re.finditer(r"\bint\s+(turtle_[a-z0-9_]+)\s*\(([^;]*?void \*stream[^;]*?)\)\s*;",
                                                           hdr, flags=re.S)))
    assert len(launching) >= 25
    for n in launching:
        op = getattr(torch.ops.turtle_b200, n[len("turtle_"):])
        assert op is not None
    # non-const pointers are declared mutated, const ones are not; the stream is not part of the schema
    s = ops.SCHEMAS["layernorm"]
    assert "Tensor(a!)? y" in s and "Tensor? x, int x_off" in s and "stream" not in s
    assert "Tensor(a!) out" in ops.SCHEMAS["gemm"] and "Tensor[] A" in ops.SCHEMAS["gemm"]
    assert "int[] y0" in ops.SCHEMAS["tile_blend"]


def test_custom_ops_trace_with_fake_tensors_and_reject_cpu_tensors():
    import pytest
    import torch
    from torch._subclasses.fake_tensor import FakeTensorMode
    from turtlevsr_b200 import ops
    with FakeTensorMode():
        x = torch.empty(100, 64, device="cuda")
        w = torch.empty(64, device="cuda")
        y = torch.empty(100, 64, device="cuda")
        assert torch.ops.turtle_b200.layernorm(x, 0, 64, w, 0, w, 0, y, 0, 64, 100, 64, 0) is None
    with pytest.raises((NotImplementedError, RuntimeError)):        # no CPU kernel is registered: there is no fallback
        torch.ops.turtle_b200.layernorm(torch.zeros(4, 64), 0, 64, torch.ones(64), 0, None, 0, torch.zeros(4, 64), 0, 64,
                                        4, 64, 0)
    # an address handed to the op layer must know its tensor
    t = torch.zeros(16)
    p = ops.devptr(t, 4)
    assert int(p) == t.data_ptr() + 16 and (p + 8).base is t and int(p + 8) == t.data_ptr() + 24
    with pytest.raises(TypeError):
        ops._split(12345)
