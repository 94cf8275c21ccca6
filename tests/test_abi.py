"""The C-ABI library loads without a GPU and exports every symbol include/vbmc_b200.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

from vbmc_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "vbmc_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vbmc_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_symbol():
    assert _lib.LIB_PATH.exists(), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.vbmc_b200_version() == 100


def test_no_cpu_fallback_without_device():
    """Without a CUDA device vbmc_b200_create must fail loudly (ENODEV), never compute on the host."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.vbmc_b200_create(ctypes.byref(h), 0)
    assert rc == _lib.ENODEV
    assert b"no CPU fallback" in lib.vbmc_b200_last_error()
    import vbmc_b200
    with pytest.raises(vbmc_b200.VbmcB200Error):
        vbmc_b200.Context(0)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under vbmc_b200/ may import or execute it."""
    for p in (ROOT / "vbmc_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
    for p in (ROOT / "vbmc_b200" / "csrc").glob("*"):
        if p.suffix in (".cu", ".cuh", ".h", ".cpp"):
            assert "oracle" not in p.read_text(), p
    # the same holds for the helper scripts and the gateways: only tests/, smoke() and bench.py's CPU-baseline / reference
    # legs may touch oracle/
    for p in list((ROOT / "tools").glob("*.py")) + list((ROOT / "mex").glob("*")):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
