"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

from conftest import ROOT


def _declared():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_builds_loads_and_exports_all_declared_symbols():
    from maggie_b200 import _build, _lib

    _build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    decl = _declared()
    assert len(decl) >= 10
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert decl == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    assert _lib.lib().mg_version() >= 100
    assert _lib.lib().mg_sites_workspace(80, 512, 512) > 0


def test_argument_errors_are_reported_not_fatal():
    from maggie_b200 import _lib

    L = _lib.lib()
    rc = L.mg_unknown_mask(None, 1, 8, 8, None, None, None, None, None)
    assert rc != 0 and b"null" in L.mg_last_error()
