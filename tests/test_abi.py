"""The C-ABI library loads without a GPU and exports every symbol include/pfs_b200.h declares
(no compute calls here: they need a device)."""
import ctypes as ct
import os
import re

import pytest

from pfs_neural_net_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "pfs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pfs_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_abi.LIB_PATH):
        _abi.build_library()
    return _abi.load_library()


def test_header_and_binding_declare_the_same_symbols():
    assert header_functions() == sorted(_abi.SYMBOLS)


def test_every_declared_symbol_is_exported(lib):
    raw = ct.CDLL(_abi.LIB_PATH)
    for name in header_functions():
        assert hasattr(raw, name), name


def test_struct_layouts_match(lib):
    assert lib.pfs_abi_version() == _abi.ABI_VERSION
    for fn, struct in _abi._SIZEOF_CHECKS.items():
        assert getattr(lib, fn)() == ct.sizeof(struct), fn


def test_supported_feature_widths(lib):
    assert lib.pfs_supports_fdim(10) == 1      # reference src/config.py:23
    assert lib.pfs_supports_fdim(16) == 1      # GNN() default, reference src/gnn.py:266
    assert lib.pfs_supports_fdim(7) == 0


def test_argument_errors_are_reported_not_thrown(lib):
    a = _abi.EdgeArgs()                        # all-NULL arguments: rejected before any CUDA call
    assert lib.pfs_edge_fwd(ct.byref(a)) == -1
    assert b"null pointer" in lib.pfs_last_error()
    t = _abi.TopologyStruct()
    t.layout, t.G, t.F, t.S, t.T, t.E = 0, 2, 10, 2394, 12, 2394 * 12
    assert lib.pfs_workspace_bytes(ct.byref(t)) > 0


def test_library_is_built_for_sm_100a():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
