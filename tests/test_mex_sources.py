"""The MEX gateways (mex/*.cpp) cannot be built here (no MATLAB, no mex.h).  They are type-checked against the C ABI
(include/vbmc_b200.h) with a declaration-only stand-in for mex.h (tests/stubs/mex.h) so that they cannot drift from the
library's signatures, and every C-ABI symbol they call must be exported by the built library."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SOURCES = sorted((ROOT / "mex").glob("*.cpp"))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
@pytest.mark.parametrize("src", SOURCES, ids=lambda p: p.name)
def test_gateway_type_checks_against_the_c_abi(src):
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'tests' / 'stubs'}",
                        f"-I{ROOT / 'include'}", f"-I{ROOT / 'mex'}", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_gateways_cover_the_shadowed_functions_and_call_exported_symbols():
    from vbmc_b200 import _lib
    names = {p.stem for p in SOURCES}
    assert {"negelcbo_vbmc_mex", "entmc_vbmc_mex", "gplogjoint_mex", "gplite_nlZ_mex", "gplite_pred_mex", "gplite_post_core_mex",
            "fminadam_negelcbo_mex"} <= names
    called = set()
    for p in list(SOURCES) + [ROOT / "mex" / "vbmc_b200_mex_common.h"]:
        called |= set(re.findall(r"\b(vbmc_b200_[a-z0-9_]+)\s*\(", p.read_text()))
    assert called and called <= set(_lib.EXPORTS), called - set(_lib.EXPORTS)
