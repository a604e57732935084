"""CPU: the C-ABI library builds, loads and exports every symbol include/gfr_b200.h declares.
No compute calls (no GPU here)."""
import ctypes
import os
import re

from geomconsistentfr_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gfr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gfr_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_all_declared_symbols():
    build.build()
    lib = ctypes.CDLL(build.LIB)
    names = declared_symbols()
    assert "gfr_shadow_march_fwd" in names and len(names) >= 5
    for n in names:
        assert hasattr(lib, n), "libgfr_b200.so does not export %s" % n


def declared_arity():
    src = open(os.path.join(ROOT, "include", "gfr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name, args in re.findall(r"\b(gfr_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def test_python_binding_covers_header():
    assert set(declared_symbols()) == set(_lib.exported_symbols())


def test_python_binding_arity_matches_header():
    """Every ctypes prototype has exactly as many arguments as the C declaration it binds."""
    arity = declared_arity()
    for name, argtypes in _lib._PROTOS.items():
        assert arity[name] == len(argtypes), (name, arity[name], len(argtypes))


def test_argument_errors_without_gpu():
    lib = _lib.load()
    assert lib.gfr_version() >= 100
    assert lib.gfr_error_string(0) == b"ok"
    # NULL pointers are rejected before any CUDA call
    assert lib.gfr_mask_pack(None, 0, 1, 256, 256, None, None) == -1
    assert lib.gfr_shadow_march_fwd(None, None, 0, None, None, 160, 0.0, None, None, None, None, None, 1, 256, 256, 1, 0, None) == -1
    assert lib.gfr_shadow_march_bwd(None, None, None, None, None, 160, None, None, 1, 256, 256, None) == -1
    assert lib.gfr_conv_tc_pack_size(16, 16, 16) == 2 * 9 * 4 * 16 * 4 and lib.gfr_conv_tc_pack_size(16, 16, 24) < 0
