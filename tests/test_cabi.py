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


def test_struct_layouts_match_the_header(tmp_path):
    """Every ctypes mirror has the size and field offsets a C compiler gives the struct declared in the header."""
    import subprocess

    from maggie_b200 import dense, optim, sparse, weights

    pairs = {"mg_wprep_layer": weights.WprepLayer, "mg_conv_desc": dense.ConvDesc, "mg_wgrad_desc": dense.WgradDesc,
             "mg_sparse_conv_desc": sparse.SparseConvDesc, "mg_optim_tensor": optim._OptimTensor, "mg_xchg_desc": dense.XchgDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "maggie_b200.h"', "int main(void) {"]
    for cname, mirror in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in mirror._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, mirror in pairs.items():
        assert int(got[cname]) == ctypes.sizeof(mirror), f"sizeof({cname})"
        for fname, _ in mirror._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(mirror, fname).offset, f"offsetof({cname}, {fname})"
