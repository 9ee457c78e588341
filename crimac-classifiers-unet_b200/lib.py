"""ctypes binding of libcrimac_b200.so — the C-ABI declared in include/crimac_b200.h.

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcrimac_b200.so")

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_float_p = ctypes.c_void_p  # device pointers are passed as integers


class CrimacError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle. Raises if the native library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CrimacError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (or crimac-classifiers-unet_b200/build.py). "
            "There is no CPU or PyTorch fallback for the U-Net hot path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.crimac_last_error.restype = ctypes.c_char_p
    lib.crimac_last_error.argtypes = []
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().crimac_last_error().decode("utf-8", "replace")
        raise CrimacError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device/host address of a torch tensor (or None) as a c_void_p."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)
