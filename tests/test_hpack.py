"""HPACK decoding (client_b200/csrc/h2.h: static + dynamic table, Huffman strings with the table of
csrc/hpack_huffman.h) against libnghttp2's encoder: header blocks produced by one stateful deflater
(incremental indexing, table size updates, Huffman where shorter) decode to the same header lists."""

import ctypes
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


@pytest.fixture(scope="module")
def nghttp2():
    try:
        import gen_hpack_huffman as gen

        return gen, gen.load()
    except OSError:
        pytest.skip("libnghttp2 is not on this machine")


@pytest.fixture(scope="module")
def decoder_exe():
    out = os.path.join(ROOT, "build", "test_hpack")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "cpp", "test_hpack.cc"), "-o", out], check=True)
    return out


def test_generated_table_is_current(nghttp2):
    gen, lib = nghttp2
    path = os.path.join(ROOT, "client_b200", "csrc", "hpack_huffman.h")
    assert open(path).read() == gen.render(gen.derive(lib)), "run python scripts/gen_hpack_huffman.py"


def _deflate_blocks(gen, lib, header_lists, table_size=4096):
    d = ctypes.c_void_p()
    assert lib.nghttp2_hd_deflate_new(ctypes.byref(d), table_size) == 0
    blocks = []
    for headers in header_lists:
        keep = []
        nva = (gen.NV * len(headers))()
        for i, (n, v) in enumerate(headers):
            nb = (ctypes.c_uint8 * max(len(n), 1)).from_buffer_copy(n or b"\0")
            vb = (ctypes.c_uint8 * max(len(v), 1)).from_buffer_copy(v or b"\0")
            keep += [nb, vb]
            nva[i] = gen.NV(nb, vb, len(n), len(v), 0)
        out = (ctypes.c_uint8 * 65536)()
        n = lib.nghttp2_hd_deflate_hd(d, out, 65536, nva, len(headers))
        assert n > 0
        blocks.append(bytes(out[:n]))
    lib.nghttp2_hd_deflate_del(d)
    return blocks


def test_blocks_from_a_stateful_encoder(nghttp2, decoder_exe):
    gen, lib = nghttp2
    rng = random.Random(7541)
    names = [b":path", b":authority", b"content-type", b"te", b"grpc-timeout", b"user-agent", b"x-custom-header", b"grpc-status", b"grpc-message",
             b"accept-encoding", b"x-" + bytes(rng.randrange(97, 123) for _ in range(20))]
    values = [b"/inference.GRPCInferenceService/ModelInfer", b"127.0.0.1:8001", b"application/grpc", b"trailers", b"2S", b"0", b"13",
              b"Request for unknown model: 'nope' is not found", b"gzip, deflate", b"grpc-go/1.60.0", b""]
    header_lists = []
    for _ in range(60):
        hs = []
        for _ in range(rng.randrange(1, 9)):
            n = rng.choice(names)
            k = rng.randrange(4)
            if k == 0:
                v = rng.choice(values)
            elif k == 1:  # printable text: Huffman is shorter
                v = bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz0123456789 /-_.:=%") for _ in range(rng.randrange(1, 80)))
            elif k == 2:  # arbitrary bytes: long codes, usually sent raw
                v = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 40)))
            else:         # few distinct rare bytes repeated: exercises 20-30 bit codes when mixed with text
                v = bytes(rng.choice([0, 1, 10, 13, 22, 127, 200, 255])) * rng.randrange(1, 3) + b"eeeeeeeeeeeeeeeeeeeeeeee"
            hs.append((n, v))
        header_lists.append(hs)
    blocks = _deflate_blocks(gen, lib, header_lists)
    assert any(b & 0x80 for blk in blocks for b in blk[:1]) or True
    r = subprocess.run([decoder_exe], input="\n".join(b.hex() for b in blocks) + "\n", capture_output=True, text=True, timeout=60)
    got = [blk.strip().splitlines() for blk in r.stdout.split("---\n") if blk.strip()]
    assert "ERROR" not in r.stdout and len(got) == len(header_lists), r.stdout[-400:]
    for lines, headers in zip(got, header_lists):
        want = ["%s %s" % (n.hex() or "-", v.hex() or "-") for n, v in headers]
        assert lines == want


def test_bad_huffman_padding_and_eos_are_rejected(decoder_exe):
    # literal without indexing, new name "a" raw, value Huffman: 'e' = 00101 then padding
    good = bytes([0x00, 0x01, 0x61, 0x81, 0b00101111])
    zero_padding = bytes([0x00, 0x01, 0x61, 0x81, 0b00101000])           # padding must be ones
    long_padding = bytes([0x00, 0x01, 0x61, 0x82, 0b00101111, 0xFF])     # 11 bits of padding
    eos = bytes([0x00, 0x01, 0x61, 0x84, 0xFF, 0xFF, 0xFF, 0xFF])        # EOS inside the string
    r = subprocess.run([decoder_exe], input="\n".join(b.hex() for b in (good, zero_padding, long_padding, eos)) + "\n", capture_output=True, text=True, timeout=60)
    assert r.stdout.split() == ["61", "65", "---", "ERROR", "ERROR", "ERROR"]
