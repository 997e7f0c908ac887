"""No-GPU checks of the drop-in boundary: the shared object builds, loads and exports every symbol that
include/ldso_b200.h declares; the product path fails loudly without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

from ldso_b200 import build as lbuild
from ldso_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "ldso_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ldso_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    lib = lbuild.build()
    assert os.path.exists(lib)
    L = capi.load()
    names = _header_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ldso_b200.h but not exported"
    # and the Python binding knows every one of them
    assert sorted(capi.SYMBOLS) == names


def test_default_settings_are_the_reference_defaults():
    s = capi.default_settings()
    assert s.huberTH == 9 and s.outlierTHSumComponent == 2500
    assert s.affineOptModeA == pytest.approx(1e12) and s.affineOptModeB == pytest.approx(1e8)
    assert s.frameEnergyTHN == pytest.approx(0.7) and s.coarseCutoffTH == 20
    assert s.solverModeDelta == pytest.approx(1e-5) and s.margWeightFac == pytest.approx(0.25)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.Error):
        capi.Context(640, 480, 4)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ldso_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in txt and "liboracle" not in txt and '#include "../oracle' not in txt, f


def test_ctypes_structs_match_the_header(tmp_path):
    """ABI drift guard: size and every field offset of the structs in include/ldso_b200.h, as a C compiler lays them out, against the
    ctypes mirrors in ldso_b200/capi.py (field names are taken from the ctypes side; a renamed or reordered field fails to compile or
    to match)."""
    import ctypes as C
    import subprocess
    pairs = [("ldso_b200_settings", capi.Settings), ("ldso_b200_window", capi.WindowC), ("ldso_b200_frame_state", capi.FrameStateC),
             ("ldso_b200_fused_io", capi.FusedIOC), ("ldso_b200_immature", capi.ImmatureC)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ldso_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, *_ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == C.sizeof(cls), (cname, got[cname], C.sizeof(cls))
        for fname, *_ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)
