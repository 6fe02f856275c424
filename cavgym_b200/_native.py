"""Loader of the CUDA extension libcavgym_sm100.so (C-ABI, include/cavgym.h).

There is no CPU implementation behind this module: if the library is missing or cannot be
loaded the import of anything that steps an environment fails with an explicit error.
"""
import ctypes
import os

from . import _abi

LIB_NAME = "libcavgym_sm100.so"
LIB_PATH = os.environ.get("CAVGYM_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)  # CAVGYM_LIB: tuning builds
_lib = None


class CavgymError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"cavgym error {code}: {message}")
        self.code = code


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(f"{LIB_NAME} is not built: run `python -m cavgym_b200.csrc.build` (nvcc, sm_100a). "
                              "cavgym_b200 has no CPU fallback.")
        _lib = _abi.bind(ctypes.CDLL(LIB_PATH))
    return _lib


def check(code):
    if code != 0:
        raise CavgymError(code, load().cavgym_last_error().decode())
    return code
