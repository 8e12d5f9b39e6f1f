"""ctypes binding of libawr_b200.so (declared in include/awr_b200.h).

There is deliberately no fallback: if the shared library is missing or a call returns
non-zero, this raises.  The product path never routes through torch ops or any CPU restatement.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# AWR_B200_DETERMINISTIC=1 selects the bit-reproducible build (make DET=1: order-independent accumulators for every sum several CTAs
# share; ~3 % slower); AWR_B200_LIB overrides the path outright (debug builds, make PROFILE=1)
LIB_PATH = os.environ.get("AWR_B200_LIB") or os.path.join(
    _HERE, "libawr_b200_det.so" if os.environ.get("AWR_B200_DETERMINISTIC") == "1" else "libawr_b200.so")

F32, BF16 = 0, 1
HUBER_MAX_BLOCKS = 1184

_HEADER = os.path.join(os.path.dirname(_HERE), "include", "awr_b200.h")
_CTYPES = {"int": C.c_int, "float": C.c_float, "long long": C.c_longlong, "unsigned": C.c_uint}


def _parse_header(path):
    """Derive ctypes prototypes from the `int awr_*(...)` declarations of include/awr_b200.h (single source of truth)."""
    import re
    src = re.sub(r"/\*.*?\*/", "", open(path).read(), flags=re.S)
    protos = {}
    for name, args in re.findall(r"\bint\s+(awr_\w+)\s*\(([^)]*)\)\s*;", src):
        types = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    types.append(C.c_void_p)
                else:
                    base = " ".join(a.replace("const", "").split()[:-1])
                    types.append(_CTYPES[base])
        protos[name] = types
    return protos


_PROTOS = _parse_header(_HEADER)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C awr-adaptive-weighting-regression_b200`). There is no CPU/PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, args in _PROTOS.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        _lib = l
    return _lib


def exported_symbols():
    return list(_PROTOS)


# ---- order-independent accumulators (awr_acc_t in include/awr_b200.h): value = hi * 2^-24 + lo * 2^-72 ---------------------
def acc_zeros(n, device):
    """n zero-initialised awr_acc_t (BatchNorm `sums` / `dsums` buffers of the C ABI)."""
    return torch.zeros(n, 2, dtype=torch.int64, device=device)


def deterministic():
    """True when the loaded library is the bit-reproducible build (slots hold two-limb integers; the default build keeps one fp32
    sum in a slot's first four bytes)."""
    return bool(lib().awr_deterministic())


def acc_to_float(t):
    """awr_acc_t[n] -> float64[n]."""
    t = t.contiguous().view(-1, 2)
    if deterministic():
        return t[:, 0].double() * 2.0 ** -24 + t[:, 1].double() * 2.0 ** -72
    return t.view(torch.float32).view(-1, 4)[:, 0].double()


def acc_from_float(v):
    """float[n] -> awr_acc_t[n] (for callers that computed the sums themselves)."""
    v = v.double().flatten()
    if deterministic():
        hi = torch.round(v * 2.0 ** 24)
        lo = torch.round((v - hi * 2.0 ** -24) * 2.0 ** 72)
        return torch.stack([hi.long(), lo.long()], dim=1).contiguous()
    out = torch.zeros(v.numel(), 4, dtype=torch.float32, device=v.device)
    out[:, 0] = v.float()
    return out.view(torch.int64).view(-1, 2).contiguous()


def check(rc: int, what: str):
    if rc != 0:
        if rc > 0:
            raise RuntimeError(f"{what}: CUDA error {rc} ({torch.cuda.get_device_name() if torch.cuda.is_available() else 'no device'})")
        raise ValueError(f"{what}: invalid argument (code {rc})")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("awr_b200 runs on CUDA tensors only (sm_100a kernels; no CPU fallback)")
