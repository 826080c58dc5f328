"""ctypes wrapper over oracle/libqvrcnn_oracle.so (the C restatement of forward_blu).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libqvrcnn_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "qvrcnn_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.qvo_model_from_bytes.restype = C.c_void_p
        L.qvo_model_from_bytes.argtypes = [C.c_char_p, C.c_size_t]
        L.qvo_model_free.argtypes = [C.c_void_p]
        L.qvo_model_file_size.restype = C.c_size_t
        L.qvo_forward_blu.restype = C.c_int
        L.qvo_forward_blu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.qvo_forward_frame_taps.restype = C.c_int
        L.qvo_forward_frame_taps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.qvo_psnr.restype = C.c_double
        L.qvo_psnr.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int64)]
        L.qvo_num_threads.restype = C.c_int
        L.qvo_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


class OracleModel:
    """Model parsed from the bytes of a static NCHW_VECT_C model file."""

    def __init__(self, model_file_bytes: bytes):
        self._h = lib().qvo_model_from_bytes(model_file_bytes, len(model_file_bytes))
        if not self._h:
            raise ValueError("oracle: bad model file image (%d bytes)" % len(model_file_bytes))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qvo_model_free(self._h)
            self._h = None

    def forward_blu(self, frames_u8: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(frames_u8, dtype=np.uint8)
        assert x.ndim == 3
        out = np.empty_like(x)
        rc = lib().qvo_forward_blu(self._h, x.ctypes.data, out.ctypes.data, x.shape[0], x.shape[1], x.shape[2])
        assert rc == 0
        return out

    def forward_taps(self, frame_u8: np.ndarray):
        """One frame -> (rec, a1[64,H,W], a2[48,H,W], a3[48,H,W], u4[H,W] int32)."""
        x = np.ascontiguousarray(frame_u8, dtype=np.uint8)
        assert x.ndim == 2
        h, w = x.shape
        rec = np.empty_like(x)
        a1 = np.empty((64, h, w), np.int8)
        a2 = np.empty((48, h, w), np.int8)
        a3 = np.empty((48, h, w), np.int8)
        u4 = np.empty((h, w), np.int32)
        rc = lib().qvo_forward_frame_taps(self._h, x.ctypes.data, rec.ctypes.data, h, w,
                                          a1.ctypes.data, a2.ctypes.data, a3.ctypes.data, u4.ctypes.data)
        assert rc == 0
        return rec, a1, a2, a3, u4


def psnr(data: np.ndarray, ori: np.ndarray):
    d = np.ascontiguousarray(data, dtype=np.uint8)
    o = np.ascontiguousarray(ori, dtype=np.uint8)
    sse = C.c_int64(0)
    p = lib().qvo_psnr(d.ctypes.data, o.ctypes.data, d.size, C.byref(sse))
    return p, sse.value


def num_threads() -> int:
    return lib().qvo_num_threads()


def set_num_threads(n: int) -> None:
    lib().qvo_set_num_threads(n)
