"""The C-ABI shared library builds for sm_100a, loads, and exports every symbol the header declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "regen_sm100.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(regen_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    names = _declared_symbols()
    assert "regen_p_sample_update" in names and "regen_denoise" in names
    for n in names:
        assert hasattr(built_lib, n), "library does not export %s" % n


def test_python_binding_covers_header(built_lib):
    from regennet_b200 import _lib
    assert sorted(_lib.exported_symbols()) == _declared_symbols()


def test_version_and_error_strings(built_lib):
    assert b"sm_100a" in built_lib.regen_version()
    assert isinstance(built_lib.regen_last_error(), bytes)


def test_argument_validation_without_gpu(built_lib):
    # null pointers / bad sizes are rejected before any CUDA call
    rc = built_lib.regen_p_sample_update(None, None, None, None, None, None, None, None, None, 16, 4, 1, 0, None)
    assert rc == -1
    assert b"null" in built_lib.regen_last_error()
    rc = built_lib.regen_rot6d_to_matrix(None, None, -5, None)
    assert rc == -1
    assert built_lib.regen_rot6d_to_matrix(None, None, 0, None) == 0  # empty input is a no-op


def test_library_contains_sm100a_code(built_lib):
    from regennet_b200 import _lib
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from regennet_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()
