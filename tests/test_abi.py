"""The C-ABI library loads and exports every symbol include/shiftgcn_b200.h declares; the ctypes mirrors of the
parameter blocks have the C compiler's sizes and offsets.  No compute calls (no GPU needed)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "shiftgcn_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as entry
    from shiftgcn_b200 import _lib
    if not os.path.exists(_lib.library_path()):
        entry.build()
    return _lib.load()


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgcn_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from shiftgcn_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/shiftgcn_b200.h but not exported"
    bound = set(_lib.SIGNATURES) | {"sgcn_last_error"}
    assert set(names) == bound, f"header and ctypes binding drifted: {set(names) ^ bound}"


def test_abi_version_and_error_channel(lib):
    assert lib.sgcn_abi_version() == 1
    assert isinstance(lib.sgcn_last_error(), bytes)


def test_struct_layouts_match_the_c_compiler(tmp_path):
    from shiftgcn_b200 import _lib
    structs = ["SgcnRowGemm", "SgcnWgrad", "SgcnTShift", "SgcnTShiftBwd", "SgcnTShiftInBwd", "SgcnTShiftInSums", "SgcnStem",
               "SgcnSideFold", "SgcnSideBwd"]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', 'int main(void) {']
    for s in structs:
        cls = getattr(_lib, s)
        lines.append(f'  printf("{s} %zu\\n", sizeof({s}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{s}.{fname} %zu\\n", offsetof({s}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    want = dict(l.split() for l in out if l.strip())
    for s in structs:
        cls = getattr(_lib, s)
        assert ctypes.sizeof(cls) == int(want[s]), s
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == int(want[f"{s}.{fname}"]), f"{s}.{fname}"


def test_no_cpu_fallback():
    """CPU tensors are refused loudly (no oracle / library fallback in the product path)"""
    import torch
    from shiftgcn_b200.modules import Shift_gcn, Shift_tcn, TCN_GCN_unit
    from shiftgcn_b200.shift import ShiftFunction
    x = torch.randn(1, 64, 4, 25)
    for mod in (Shift_gcn(64, 64, None).cpu(), Shift_tcn(64, 64).cpu(), TCN_GCN_unit(64, 64, None).cpu()):
        with pytest.raises(RuntimeError):
            mod(x)
    with pytest.raises(RuntimeError):
        ShiftFunction.apply(x, torch.zeros(64), torch.zeros(64), 1)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "shiftgcn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", ""), f"{f} references oracle/"


def test_drop_in_import_paths():
    """the reference's own import statements resolve (model/shift_gcn.py:9-11, main.py:256)"""
    code = ("import sys; sys.path.append('./model/Temporal_shift/'); from cuda.shift import Shift, ShiftFunction; "
            "import model.shift_gcn as m; assert m.Shift is Shift; "
            "[getattr(m, n) for n in ('Model','Shift_gcn','Shift_tcn','TCN_GCN_unit','tcn','import_class','conv_init','bn_init')]")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)
