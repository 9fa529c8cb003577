"""Loads libvelox_b200.so (built in-tree by `make` / __graft_entry__.build()).

There is no CPU fallback: if the library is missing or a kernel fails, the caller gets an
exception (VeloxRuntimeError / VeloxUserError analogues below)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libvelox_b200.so")

VB2_OK, VB2_ERR_CUDA, VB2_ERR_INVALID, VB2_ERR_UNSUPPORTED, VB2_ERR_USER = 0, 1, 2, 3, 4


class VeloxRuntimeError(RuntimeError):
    """Mirrors facebook::velox::VeloxRuntimeError (VELOX_CHECK failures, CUDA errors)."""


class VeloxUserError(RuntimeError):
    """Mirrors facebook::velox::VeloxUserError (arithmetic overflow, division by zero, bad cast)."""


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VeloxRuntimeError(
                f"{LIB_PATH} is missing: run `make` (or __graft_entry__.build()). "
                "velox_b200 has no CPU fallback.")
        # torch first: its bundled libnccl.so.2 / libcudart are then the ones our library binds to.
        import torch  # noqa: F401
        _lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        _lib.vb2_last_error.restype = C.c_char_p
    return _lib


def check(code: int) -> None:
    if code == VB2_OK:
        return
    msg = lib().vb2_last_error().decode(errors="replace")
    if code == VB2_ERR_USER:
        raise VeloxUserError(msg)
    raise VeloxRuntimeError(f"[{code}] {msg}")
