"""ctypes front-end of oracle/lbic_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It restates the two CompressAI native functions the reference calls
(graphs/layers/entropy_layers_cai.py:61-64, graphs/models/BlockBasedImgCompLossy_net.py:328,
359-360, 409-410, 439).  PARITY UNPINNED versus a real CompressAI build (see the C header).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "lbic_oracle.c")
_SO = os.path.join(_HERE, "_build", "liblbic_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (no GPU needed)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        i32p = ctypes.POINTER(ctypes.c_int32)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.oracle_pmf_to_quantized_cdf.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_int,
                                                  ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
        L.oracle_pmf_to_quantized_cdf.restype = ctypes.c_int
        L.oracle_rans_encode.argtypes = [i32p, i32p, ctypes.c_long, i32p, ctypes.c_int, i32p, i32p,
                                         u8p, ctypes.c_long]
        L.oracle_rans_encode.restype = ctypes.c_long
        L.oracle_rans_decode.argtypes = [u8p, ctypes.c_long, i32p, ctypes.c_long, i32p, ctypes.c_int,
                                         i32p, i32p, i32p]
        L.oracle_rans_decode.restype = ctypes.c_int
        L.oracle_rans_dec_init.argtypes = [ctypes.c_void_p, u8p, ctypes.c_long]
        L.oracle_rans_dec_init.restype = None
        L.oracle_rans_dec_stream.argtypes = [ctypes.c_void_p, i32p, ctypes.c_long, i32p, ctypes.c_int,
                                             i32p, i32p, i32p]
        L.oracle_rans_dec_stream.restype = ctypes.c_int
        L.oracle_rans_dec_consumed.argtypes = [ctypes.c_void_p, u8p]
        L.oracle_rans_dec_consumed.restype = ctypes.c_long
        _lib = L
    return _lib


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty))


def pmf_to_quantized_cdf(pmf, precision: int = 16) -> np.ndarray:
    """compressai._CXX.pmf_to_quantized_cdf(pmf: list[float], precision) -> list[int]."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(p.size + 1, dtype=np.uint32)
    rc = lib().oracle_pmf_to_quantized_cdf(_p(p, ctypes.c_float), p.size, precision,
                                           _p(out, ctypes.c_uint32))
    if rc != 0:
        raise ValueError(f"pmf_to_quantized_cdf failed rc={rc}")
    return out.astype(np.int64)


class Tables:
    """(quantized_cdf [T, L] int32, cdf_length [T] int32, offset [T] int32), as the reference
    passes them to the coder (BlockBasedImgCompLossy_net.py:322-324)."""

    def __init__(self, cdf, cdf_length, offset):
        self.cdf = _i32(cdf)
        assert self.cdf.ndim == 2
        self.cdf_length = _i32(cdf_length)
        self.offset = _i32(offset)

    def args(self):
        return (_p(self.cdf, ctypes.c_int32), int(self.cdf.shape[1]),
                _p(self.cdf_length, ctypes.c_int32), _p(self.offset, ctypes.c_int32))


def rans_encode(symbols, indexes, tables: Tables) -> bytes:
    """BufferedRansEncoder.encode_with_indexes(...) followed by flush()."""
    s, ix = _i32(symbols).ravel(), _i32(indexes).ravel()
    assert s.size == ix.size
    cap = 4 * (2 * s.size + 64)
    out = np.empty(cap, dtype=np.uint8)
    n = lib().oracle_rans_encode(_p(s, ctypes.c_int32), _p(ix, ctypes.c_int32), s.size,
                                 *tables.args(), _p(out, ctypes.c_uint8), cap)
    if n < 0:
        raise RuntimeError(f"oracle_rans_encode failed rc={n}")
    return out[:n].tobytes()


def rans_decode(stream: bytes, indexes, tables: Tables) -> np.ndarray:
    """RansDecoder.set_stream + decode_stream over all indexes at once."""
    ix = _i32(indexes).ravel()
    buf = np.frombuffer(stream, dtype=np.uint8).copy()
    out = np.empty(ix.size, dtype=np.int32)
    rc = lib().oracle_rans_decode(_p(buf, ctypes.c_uint8), buf.size, _p(ix, ctypes.c_int32), ix.size,
                                  *tables.args(), _p(out, ctypes.c_int32))
    if rc != 0:
        raise RuntimeError("oracle_rans_decode failed")
    return out


class RansDecoder:
    """Stateful twin of compressai.ans.RansDecoder (set_stream / decode_stream)."""

    def __init__(self):
        self._state = ctypes.create_string_buffer(32)
        self._buf = None

    def set_stream(self, stream: bytes):
        self._buf = np.frombuffer(stream, dtype=np.uint8).copy()
        lib().oracle_rans_dec_init(self._state, _p(self._buf, ctypes.c_uint8), self._buf.size)

    def decode_stream(self, indexes, tables: Tables) -> np.ndarray:
        ix = _i32(indexes).ravel()
        out = np.empty(ix.size, dtype=np.int32)
        lib().oracle_rans_dec_stream(self._state, _p(ix, ctypes.c_int32), ix.size, *tables.args(),
                                     _p(out, ctypes.c_int32))
        return out

    def consumed(self) -> int:
        return int(lib().oracle_rans_dec_consumed(self._state, _p(self._buf, ctypes.c_uint8)))
