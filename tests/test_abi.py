"""CPU: the C-ABI library builds, loads, and exports exactly what include/lyssa_b200.h
declares; the Python binding table covers every declaration.  No compute calls here."""
import ctypes
import os
import re

import pytest

from lyssandra_b200 import _native, _build

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "lyssa_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lys_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "liblyssa_b200.so lacks %s declared in include/lyssa_b200.h" % name
    assert sorted(_native.SIGNATURES) == names, "binding table and header disagree"


def test_library_is_in_tree_and_sm100a():
    path = _native.lib_path()
    assert os.path.isfile(path) and path.startswith(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert "compute_100a" in " ".join(_build.NVCC_FLAGS)


def test_pure_host_entry_points():
    lib = _native.load()
    assert lib.lys_version() >= 100
    # workspace queries are host arithmetic only
    assert lib.lys_bomp_workspace_bytes(64, 1024, 1 << 20, 5) >= 1024 * 1024 * 4
    assert lib.lys_bomp_workspace_bytes(0, 1024, 10, 5) == 0
    assert lib.lys_residual_workspace_bytes(64, 1024, 1000) >= 64 * 1024 * 4
    assert lib.lys_ksvd_sweep_workspace_bytes(64, 1024, 1000, 5) >= 64 * 1024 * 4
    assert lib.lys_odl_update_workspace_bytes(128, 2048) >= 128 * 2048 * 4 and lib.lys_odl_accumulate_workspace_bytes(2048, 4096, 5) >= 4096 * 5 * 4
    # argument validation happens before any CUDA call
    rc = lib.lys_bomp_encode(None, 1, 64, None, 1024, None, 64, 1024, 10, 0, None, None, None, None, 1, 1024, None, 0, None)
    assert rc == _native.LYS_EINVAL
    assert b"n_nonzero_coefs" in lib.lys_last_error()
    rc = lib.lys_bomp_encode(None, 1, 64, None, 1024, None, 64, 5000, 10, 5, None, None, None, None, 1, 5000, None, 0, None)
    assert rc == _native.LYS_EINVAL and b"K=5000" in lib.lys_last_error()
    with pytest.raises(_native.LyssaError):
        _native.check(rc)
    out = ctypes.c_void_p()
    assert lib.lys_comm_create(0, 1, ctypes.byref(out)) == 0


def test_build_fingerprint_travels_with_the_tree(tmp_path, monkeypatch):
    """the library says what it was built from, and the hash does not depend on WHERE the tree lives: the GPU box runs
    the repo from another directory, and a fingerprint over absolute paths made every rank of every launch rebuild the
    .so there (and race on the half-written file)"""
    import shutil
    lib = _native.load()
    here = _build._fingerprint()
    assert lib.lys_build_fingerprint().decode() == here
    assert _build.build() == _native.lib_path()                      # fresh: returns without compiling
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    moved = tmp_path / "elsewhere" / "lyssandra_b200"
    shutil.copytree(os.path.join(root, "lyssandra_b200", "csrc"), moved / "csrc")
    shutil.copytree(os.path.join(root, "include"), tmp_path / "elsewhere" / "include")
    monkeypatch.setattr(_build, "CSRC", str(moved / "csrc"))
    monkeypatch.setattr(_build, "_PKG", str(moved))
    assert _build._fingerprint() == here
    (moved / "csrc" / "common.cuh").write_text((moved / "csrc" / "common.cuh").read_text() + "\n// edited\n")
    assert _build._fingerprint() != here
