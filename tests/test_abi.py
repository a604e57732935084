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


def test_python_binding_covers_header():
    assert set(declared_symbols()) == set(_lib.exported_symbols())


def test_argument_errors_without_gpu():
    lib = _lib.load()
    assert lib.gfr_version() >= 100
    assert lib.gfr_error_string(0) == b"ok"
    # NULL pointers are rejected before any CUDA call
    assert lib.gfr_mask_pack(None, 0, 1, 256, 256, None, None) == -1
    assert lib.gfr_shadow_march_fwd(None, None, 0, None, None, 160, 0.0, None, None, None, 1, 256, 256, 0, None) == -1
