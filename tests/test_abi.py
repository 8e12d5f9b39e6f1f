"""CPU-side checks of the C-ABI library: it loads and exports every symbol include/awr_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "awr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(awr_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from awr_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 6
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/awr_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == names, "python binding and header disagree"
    assert lib.awr_version() == 100


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "awr-adaptive-weighting-regression_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "oracle" not in s, f"{f} references oracle/"
