"""Loader for libpepper_b200.so (the C-ABI of include/pepper_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing, or no CUDA
device is visible, the compute entry points raise."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB_LIB_PATH") or os.path.join(HERE, "libpepper_b200.so")      # PB_LIB_PATH: A/B builds


class PepperB200Error(RuntimeError):
    """`rc` carries the C-ABI return code (PB_ERR_*), so callers compare codes instead of parsing the message."""

    def __init__(self, msg: str, rc: int = 0):
        super().__init__(msg)
        self.rc = rc


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PepperB200Error(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  pepper_b200 has no CPU fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.pb_last_error.restype = C.c_char_p
        _lib.pb_version.restype = C.c_int
        _lib.pb_device_count.restype = C.c_int
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().pb_last_error().decode(errors="replace")
        raise PepperB200Error(f"{what} failed (code {rc}): {msg}", rc)


def device_count() -> int:
    return int(lib().pb_device_count())


def require_gpu() -> None:
    if device_count() < 1:
        raise PepperB200Error("no CUDA device visible: pepper_b200 runs on sm_100a GPUs only (no CPU fallback)")
