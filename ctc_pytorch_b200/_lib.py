"""ctypes binding of libctcb200.so.

The C prototypes are read from include/ctcb200.h so the header stays the single source of truth for
the ABI. There is deliberately no fallback: if the shared object is missing or a call fails, the
caller gets a RuntimeError — the product path never silently degrades to PyTorch or CPU code.
"""
import ctypes
import functools
import os
import re

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libctcb200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "ctcb200.h")

_CT = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "uint32_t": ctypes.c_uint32,
    "int32_t": ctypes.c_int32,
    "ctcb200_stream_t": ctypes.c_void_p,
}


def parse_header(path=HEADER_PATH):
    """Return {name: (restype, [argtypes], [argnames])} for every CTCB200_API prototype."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"CTCB200_API\s+([\w\s\*]+?)\s*(\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "*" in ret:
            restype = ctypes.c_char_p if "char" in ret else ctypes.c_void_p
        else:
            restype = _CT[ret.replace("const", "").strip()]
        argtypes, argnames = [], []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                if "*" in p:
                    argtypes.append(ctypes.c_void_p)
                    argnames.append(p.split("*")[-1].strip())
                else:
                    toks = p.replace("const ", "").split()
                    argtypes.append(_CT[toks[0]])
                    argnames.append(toks[-1])
        protos[name] = (restype, argtypes, argnames)
    return protos


# kernels of ours enqueued by one call of each entry point (memsets not counted)
KERNELS_PER_CALL = {"ctcb200_stream_wait_geq": 0, "ctcb200_stream_write_value": 0, "ctcb200_greedy_decode": 2, "ctcb200_bn_train_stats": 2, "ctcb200_bn_bwd": 2,
                    "ctcb200_bn_bwd_coef": 2}


class _Lib(object):
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libctcb200.so is not built (expected at %s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `python -m ctc_pytorch_b200._build`; there is no CPU / PyTorch fallback." % LIB_PATH)
        self.dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        self.launches = 0  # number of our kernels enqueued so far through the C ABI
        for name, (restype, argtypes, _) in self.protos.items():
            fn = getattr(self.dll, name)  # AttributeError here = header/library mismatch
            fn.restype = restype
            fn.argtypes = argtypes

    def last_error(self):
        return self.dll.ctcb200_last_error().decode("utf-8", "replace")

    def call(self, name, *args):
        """Invoke an int-returning entry point; raise RuntimeError with the library's message on failure."""
        rc = getattr(self.dll, name)(*args)
        self.launches += KERNELS_PER_CALL.get(name, 1)
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (name, rc, self.last_error()))
        if _DEBUG_SYNC:  # development aid: surface asynchronous kernel faults at the call that caused them
            try:
                torch.cuda.synchronize()
            except Exception as e:
                raise RuntimeError("%s: kernel fault surfaced at synchronize: %s" % (name, str(e).split("\n")[0]))
        return rc


_LIB = None
_DEBUG_SYNC = os.environ.get("CTCB200_DEBUG_SYNC", "0") == "1"


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _Lib()
    return _LIB


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def stream():
    """Current stream of the CURRENT device: every public entry point runs under `on_tensor_device`, which makes the
    tensors' device current first, so the launch, its stream and the library's per-device state always agree."""
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_tensor_device(fn):
    """Decorator: run `fn` with the device of its first CUDA tensor argument as the current CUDA device (a model on cuda:1
    while cuda:0 is current would otherwise launch on the wrong device / stream)."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        for a in list(args) + list(kw.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kw)
        return fn(*args, **kw)
    return wrapper


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("ctc_pytorch_b200 runs on CUDA tensors only (got a %s tensor); there is no CPU path"
                               % t.device)
