"""client_b200/csrc/philox.cuh (the math the kernels inline) compiled for the host
and compared with the oracle: every dtype of the fill contract, all 256 pixel values
of every scaling mode, and the software fp16 conversions.  A test aid: the host build
is never part of libtb200.so."""

import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import cref, image

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul.cc")
LIB = os.path.join(HERE, "host_emul", "libemul.so")


@pytest.fixture(scope="module")
def emul():
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    hdrs = [os.path.join(HERE, "..", "client_b200", "csrc", h) for h in ("philox.cuh", "deflate.cuh", "resample.h")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max([os.path.getmtime(SRC)] + [os.path.getmtime(h) for h in hdrs]):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC], check=True)
    L = ctypes.CDLL(LIB)
    for f in (L.emul_fill, L.emul_fill_hoisted):
        f.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                      ctypes.c_double, ctypes.c_double, ctypes.c_int64, ctypes.c_uint64]
        f.restype = None
    for f in (L.emul_scale_f32_bits, L.emul_scale_f16_bits):
        f.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, ctypes.c_int]
        f.restype = ctypes.c_uint32
    for f in (L.emul_f32_to_f16, L.emul_f16_to_f32):
        f.argtypes = [ctypes.c_uint32]
        f.restype = ctypes.c_uint32
    return L


TYPES = ["FP32", "FP16", "BF16", "FP64", "INT64", "UINT64", "INT32", "UINT32", "INT16", "UINT16", "INT8", "UINT8", "BOOL"]


@pytest.mark.parametrize("dt", TYPES)
def test_fill_contract_header_vs_oracle(emul, dt):
    is_float = dt in ("FP32", "FP16", "BF16", "FP64")
    cases = [dict()]
    if is_float:
        cases += [dict(lo=-1.0, span=2.0), dict(lo=0.5, span=1e-3), dict(lo=0.0, span=255.0)]
    elif dt != "BOOL":
        top = {"INT8": 200, "UINT8": 256, "INT16": 60000, "UINT16": 65536}.get(dt, 128256)
        cases += [dict(ilo=-5, irange=top), dict(ilo=0, irange=2), dict(ilo=7, irange=1)]
    for kw in cases:
        n = 4099
        out = np.zeros(n, np.uint8)
        emul.emul_fill(out.ctypes.data, n, cref.DT[dt], 0xABCDEF0123456789, (9 << 32) | 77,
                       kw.get("lo", 0.0), kw.get("span", 0.0), kw.get("ilo", 0), kw.get("irange", 0))
        ref = cref.fill(n, dt, seed=0xABCDEF0123456789, stream=(9 << 32) | 77, **kw)
        assert np.array_equal(out, ref), (dt, kw)


@pytest.mark.parametrize("dt", ["FP32", "FP16", "INT64", "UINT8"])
def test_stream_hoisted_philox_is_the_same_function(emul, dt):
    """philox_stream_const + philox4x32_10_hoisted (what fill_uniform_kernel runs: 18 multiplies
    per call) against the C oracle's plain ten rounds, over seeds and streams that exercise
    every carry of the folded rounds."""
    rng = np.random.default_rng(5)
    kw = dict(lo=-2.0, span=5.0) if dt in ("FP32", "FP16") else (dict(ilo=0, irange=30522) if dt == "INT64" else dict())
    for seed, stream in [(0, 0), (1, 0xFFFFFFFF), (0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF), (0xABCDEF0123456789, (9 << 32) | 77)] + [
            (int(rng.integers(0, 2**63)), int(rng.integers(0, 2**63))) for _ in range(8)]:
        n = 16 * 3001
        out = np.zeros(n, np.uint8)
        emul.emul_fill_hoisted(out.ctypes.data, n, cref.DT[dt], seed, stream, kw.get("lo", 0.0), kw.get("span", 0.0),
                               kw.get("ilo", 0), kw.get("irange", 0))
        assert np.array_equal(out, cref.fill(n, dt, seed=seed, stream=stream, **kw)), (dt, hex(seed), hex(stream))


def test_scaling_all_256_values(emul):
    for code, name in ((0, "NONE"), (1, "INCEPTION"), (2, "VGG")):
        for c in (1, 3):
            for ch in range(c):
                px = np.arange(256, dtype=np.uint8).reshape(1, 256, 1).repeat(c, axis=2)
                r32 = image.preprocess_pixels(px, np.float32, name, False)[0, :, ch].view(np.uint32)
                r16 = image.preprocess_pixels(px, np.float16, name, False)[0, :, ch].view(np.uint16)
                m32 = np.array([emul.emul_scale_f32_bits(i, code, c, ch) for i in range(256)], dtype=np.uint32)
                m16 = np.array([emul.emul_scale_f16_bits(i, code, c, ch) for i in range(256)], dtype=np.uint16)
                assert np.array_equal(r32, m32) and np.array_equal(r16, m16), (name, c, ch)


def test_software_half_conversions(emul):
    allh = np.arange(65536, dtype=np.uint16)
    want = allh.view(np.float16).astype(np.float32).view(np.uint32)
    got = np.array([emul.emul_f16_to_f32(int(h)) for h in allh], dtype=np.uint32)
    ok = ~np.isnan(allh.view(np.float16))
    assert np.array_equal(want[ok], got[ok])
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2**32, 50000, dtype=np.uint64).astype(np.uint32)
    edge = np.array([0x477FE000, 0x477FEFFF, 0x477FF000, 0x38800000, 0x387FFFFF, 0x33000000, 0x33000001, 0x32FFFFFF,
                     0x3F801000, 0x3F803000, 0x3F802000, 0x7F800000, 0xFF800000, 0, 0x80000000], dtype=np.uint32)
    bits = np.concatenate([bits, edge])
    with np.errstate(over="ignore"):
        want = bits.view(np.float32).astype(np.float16).view(np.uint16)
    got = np.array([emul.emul_f32_to_f16(int(b)) for b in bits], dtype=np.uint16)
    ok = ~np.isnan(bits.view(np.float32))
    assert np.array_equal(want[ok], got[ok])


def test_resize_tables_match_the_pillow_pinned_oracle(emul):
    """client_b200/csrc/resample.h (what the runtime uploads for the resize kernel) ==
    oracle.image.resample_coefficients, which tests/test_oracle.py pins against Pillow."""
    from oracle import image

    emul.emul_resample_tables.restype = ctypes.c_int
    emul.emul_resample_tables.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    for in_size, out_size in [(500, 224), (375, 224), (224, 224), (37, 224), (1920, 224), (4000, 224), (7, 23), (23, 7), (1, 4), (449, 224), (299, 600)]:
        want_b, want_c = image.resample_coefficients(in_size, out_size)
        cap = want_c.shape[1]
        bounds = np.zeros((out_size, 2), dtype=np.int32)
        coeffs = np.zeros((out_size, cap), dtype=np.int32)
        ks = emul.emul_resample_tables(in_size, out_size, bounds.ctypes.data, coeffs.ctypes.data, cap)
        assert ks == cap
        assert np.array_equal(bounds, want_b) and np.array_equal(coeffs, want_c), (in_size, out_size)


def test_bytes_fill_header_vs_oracle(emul):
    """fill_group_bytes (client_b200/csrc/philox.cuh) == oracle_fill for BYTES, and the result
    is what deserialize_bytes_tensor expects: count strings of the requested length."""
    from client_b200.utils import deserialize_bytes_tensor
    from oracle import cref

    emul.emul_fill_bytes.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32]
    emul.emul_fill_bytes.restype = None
    for count, length in [(1, 0), (5, 1), (7, 13), (100, 128), (3, 1000), (16, 12)]:
        n = count * (4 + length)
        got = np.zeros(max(n, 1), np.uint8)
        emul.emul_fill_bytes(got.ctypes.data, n, 77, 5, length)
        want = cref.fill(n, "BYTES", seed=77, stream=5, irange=length)
        assert np.array_equal(got[:n], want), (count, length)
        strings = deserialize_bytes_tensor(want.tobytes())
        assert len(strings) == count and all(len(x) == length and (length == 0 or x.isalnum()) for x in strings)


def _deflate_inputs():
    rng = np.random.default_rng(12)
    yield "empty", b""
    yield "one byte", b"\x07"
    yield "zeros 100k", bytes(100000)
    yield "random 20k (stored fallback)", rng.integers(0, 256, 20000, dtype=np.uint8).tobytes()
    yield "text", (b"the quick brown fox jumps over the lazy dog. " * 700)
    yield "int64 token ids", rng.integers(0, 30522, 384 * 5, dtype=np.int64).tobytes()
    yield "fp32 unit interval", rng.random(6000, dtype=np.float32).tobytes()
    yield "exactly one chunk", bytes(range(256)) * 32
    yield "chunk + 1", bytes(range(256)) * 32 + b"x"
    yield "high bytes", bytes([200 + (i % 50) for i in range(30000)])
    yield "runs of 258+", b"".join(bytes([i]) * (300 + i) for i in range(60))
    yield "int32 token ids", rng.integers(0, 128256, 4096 * 3, dtype=np.int32).tobytes()
    yield "repeated int64 tokens", np.tile(rng.integers(0, 30522, 64, dtype=np.int64), 40).tobytes()
    yield "two symbols", bytes([7, 9]) * 9000
    yield "one symbol, no matches possible", b"ab"
    yield "json header + tensor", b'{"inputs":[{"name":"input_ids","shape":[1,384],"datatype":"INT64","parameters":{"binary_data_size":3072}}]}' + rng.integers(0, 30522, 384, dtype=np.int64).tobytes()


@pytest.mark.parametrize("gzip_format", [0, 1])
def test_device_deflate_logic_round_trips_through_zlib(emul, gzip_format):
    """client_b200/csrc/deflate.cuh run on the CPU: the stream the encoder produces is decoded
    by the reference's own decompressor (zlib / gzip, PY/http/_infer_result.py:71-76 and the
    server side of PY/http/_client.py:1440-1460) back to the input, checksums included."""
    import gzip
    import zlib

    emul.emul_deflate.restype = ctypes.c_uint64
    emul.emul_deflate.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p]
    for label, data in _deflate_inputs():
        src = np.frombuffer(data, dtype=np.uint8).copy() if data else np.zeros(1, np.uint8)
        dst = np.zeros(len(data) + (len(data) // 8192 + 2) * 32 + 64, dtype=np.uint8)
        n = emul.emul_deflate(src.ctypes.data, len(data), gzip_format, dst.ctypes.data)
        stream = dst[:n].tobytes()
        back = gzip.decompress(stream) if gzip_format else zlib.decompress(stream)
        assert back == data, label
        if label in ("zeros 100k", "text", "int64 token ids", "runs of 258+"):
            assert n < len(data) * 0.6, (label, n, len(data))


def test_dynamic_huffman_reaches_zlib_level_6_on_token_ids(emul):
    """with a code per chunk the encoder is within 10 % of zlib level 6 on token ids
    (round 1, fixed codes: 0.44 against 0.34) -- measured here on the CPU emulation of the kernel's logic;
    every chunk of such data ends up as a dynamic block."""
    import zlib

    emul.emul_deflate_modes.restype = ctypes.c_uint64
    emul.emul_deflate_modes.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]
    rng = np.random.default_rng(3)
    for label, data, slack in (("INT64 [0,30522)", rng.integers(0, 30522, 1 << 16, dtype=np.int64).tobytes(), 1.10),
                               ("INT32 [0,128256)", rng.integers(0, 128256, 1 << 17, dtype=np.int32).tobytes(), 1.10),
                               ("attention mask 0/1", rng.integers(0, 2, 1 << 16, dtype=np.int64).tobytes(), 2.0)):
        src = np.frombuffer(data, dtype=np.uint8).copy()
        dst = np.zeros(len(data) * 2, dtype=np.uint8)
        modes = (ctypes.c_uint32 * 8)()
        n = emul.emul_deflate_modes(src.ctypes.data, len(data), 0, dst.ctypes.data, modes)
        assert zlib.decompress(dst[:n].tobytes()) == data
        ratio, ref = n / len(data), len(zlib.compress(data, 6)) / len(data)
        assert ratio <= ref * slack, (label, ratio, ref)
        assert modes[2] == len(data) // 8192 and modes[0] == 0, (label, list(modes))
