"""ctypes face of oracle/liboracle.so (the plain-C restatement). TEST INFRASTRUCTURE."""

import ctypes
import os

import numpy as np

from .build import LIB, build_oracle

DT = {"BOOL": 1, "UINT8": 2, "UINT16": 3, "UINT32": 4, "UINT64": 5, "INT8": 6, "INT16": 7,
      "INT32": 8, "INT64": 9, "FP16": 10, "FP32": 11, "FP64": 12, "BYTES": 13, "BF16": 14}
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build_oracle()
        L = ctypes.CDLL(LIB)
        vp, u64, u32, i64, dbl = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64, ctypes.c_double
        L.oracle_philox4x32_10.argtypes = [vp, vp, vp]
        L.oracle_philox4x32_10.restype = None
        L.oracle_fill.argtypes = [vp, u64, u32, u32, u64, u64, dbl, dbl, i64, u64]
        L.oracle_fill.restype = None
        L.oracle_pack_image.argtypes = [vp, u32, u32, vp] + [ctypes.c_int] * 4 + [u32]
        L.oracle_pack_image.restype = ctypes.c_int
        L.oracle_checksum.argtypes = [vp, u64, ctypes.POINTER(u64), ctypes.POINTER(u32)]
        L.oracle_checksum.restype = None
        L.oracle_count_diff_bytes.argtypes = [vp, vp, u64]
        L.oracle_count_diff_bytes.restype = u64
        L.oracle_addsub_mismatches.argtypes = [vp, vp, vp, vp, u64]
        L.oracle_addsub_mismatches.restype = u64
        L.oracle_top1.argtypes = [vp, u64, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(u64)]
        L.oracle_top1.restype = u32
        L.oracle_topk.argtypes = [vp, u64, u32, vp, vp]
        L.oracle_topk.restype = None
        L.oracle_marshal_http.argtypes = [vp, vp, u64, vp, vp, u64]
        L.oracle_marshal_http.restype = None
        _lib = L
    return _lib


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    lib().oracle_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def fill(nbytes, datatype, seed=0, stream=0, mode=0, lo=0.0, span=0.0, ilo=0, irange=0):
    """bytes of a generated tensor (uint8 array of length nbytes)."""
    out = np.zeros(max(int(nbytes), 1), np.uint8)
    lib().oracle_fill(out.ctypes.data, int(nbytes), DT[datatype], int(mode), int(seed) & (2**64 - 1),
                      int(stream) & (2**64 - 1), float(lo), float(span), int(ilo), int(irange))
    return out[: int(nbytes)]


def pack_image(images_u8_nhwc, datatype, layout="NCHW", scaling="NONE"):
    src = np.ascontiguousarray(images_u8_nhwc, dtype=np.uint8)
    n, h, w, c = src.shape
    es = 4 if datatype == "FP32" else 2
    out = np.zeros(n * h * w * c * es, np.uint8)
    rc = lib().oracle_pack_image(out.ctypes.data, DT[datatype], 0 if layout == "NCHW" else 1,
                                 src.ctypes.data, n, h, w, c, {"NONE": 0, "INCEPTION": 1, "VGG": 2}[scaling])
    assert rc == 0
    return out


def checksum(buf):
    a = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    s, x = ctypes.c_uint64(), ctypes.c_uint32()
    lib().oracle_checksum(a.ctypes.data, a.size, ctypes.byref(s), ctypes.byref(x))
    return int(s.value), int(x.value)


def count_diff_bytes(a, b):
    a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    b = np.ascontiguousarray(b).view(np.uint8).reshape(-1)
    return int(lib().oracle_count_diff_bytes(a.ctypes.data, b.ctypes.data, min(a.size, b.size)))


def addsub_mismatches(out0, out1, in0, in1):
    arrs = [np.ascontiguousarray(x, dtype=np.int32).reshape(-1) for x in (out0, out1, in0, in1)]
    return int(lib().oracle_addsub_mismatches(*[a.ctypes.data for a in arrs], arrs[0].size))


def top1(values):
    v = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
    mv, bad = ctypes.c_float(), ctypes.c_uint64()
    idx = lib().oracle_top1(v.ctypes.data, v.size, ctypes.byref(mv), ctypes.byref(bad))
    return int(idx), float(mv.value), int(bad.value)


def topk(values, k):
    """(values float32[k], indices uint32[k]) in numpy.argsort(-x, kind="stable") order."""
    v = np.ascontiguousarray(values, dtype=np.float32).reshape(-1)
    vals = np.zeros(k, dtype=np.float32)
    idx = np.zeros(k, dtype=np.uint32)
    lib().oracle_topk(v.ctypes.data, v.size, int(k), vals.ctypes.data, idx.ctypes.data)
    return vals, idx

