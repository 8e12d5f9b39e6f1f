"""ctypes loader of libawr_b200_debug.so (hardware probes; `make -C awr-adaptive-weighting-regression_b200 debug`), prototypes from
include/awr_b200_debug.h.  Debug tooling only: nothing under the product package imports this."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from awr_b200 import _lib as L          # noqa: E402  (header parser, error helper)

_dbg = None


def lib():
    global _dbg
    if _dbg is None:
        path = os.path.join(ROOT, "awr-adaptive-weighting-regression_b200", "libawr_b200_debug.so")
        d = C.CDLL(path)
        for name, args in L._parse_header(os.path.join(ROOT, "include", "awr_b200_debug.h")).items():
            fn = getattr(d, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _dbg = d
    return _dbg
