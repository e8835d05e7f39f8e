"""The oracle against everything that pins it (CPU only):
  * Random123 known-answer vectors for Philox4x32-10 and cuRAND's host generator
  * the C restatement against the numpy restatement
  * golden wire vectors generated from the reference client (tests/golden/wire_golden.json)
  * fixtures from the reference's image_client.preprocess (tests/golden/image_golden.npz)
  * the reference's own known-answer tests quoted in SURVEY.md 8c / 9.4
"""

import ctypes
import glob
import json
import os

import numpy as np
import pytest

from oracle import cref, fill, image, wire

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = [  # Random123 kat_vectors, philox4x32 10 rounds: ctr, key, expected
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0], [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLD, "wire_golden.json")) as fh:
        return json.load(fh)["cases"]


def test_philox_known_answers():
    for ctr, key, want in KAT:
        assert list(cref.philox(ctr, key)) == want
        got = fill.philox4x32_10(*[np.array([c], np.uint64) for c in ctr], key[0], key[1])
        assert [int(x[0]) for x in got] == want


def test_philox_matches_curand_host_generator():
    """cuRAND's CPU generator (no GPU needed): first block of the sequence for a
    seed is philox(ctr=0, key=seed)."""
    libs = sorted(glob.glob("/usr/local/cuda/lib64/libcurand.so.*"))
    if not libs:
        pytest.skip("libcurand not present")
    cu = ctypes.CDLL(libs[0])
    for seed in (0, 1, 0x123456789ABCDEF, 2**64 - 1):
        gen = ctypes.c_void_p()
        assert cu.curandCreateGeneratorHost(ctypes.byref(gen), 161) == 0  # CURAND_RNG_PSEUDO_PHILOX4_32_10
        assert cu.curandSetPseudoRandomGeneratorSeed(gen, ctypes.c_ulonglong(seed)) == 0
        out = np.zeros(4, np.uint32)
        assert cu.curandGenerate(gen, out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(4)) == 0
        cu.curandDestroyGenerator(gen)
        assert list(out) == list(cref.philox([0, 0, 0, 0], [seed & 0xFFFFFFFF, seed >> 32]))


ALL_TYPES = ["FP32", "FP16", "BF16", "FP64", "INT64", "UINT64", "INT32", "UINT32", "INT16", "UINT16", "INT8", "UINT8", "BOOL"]


@pytest.mark.parametrize("dt", ALL_TYPES)
def test_c_and_numpy_fill_oracles_agree(dt):
    is_float = dt in ("FP32", "FP16", "BF16", "FP64")
    cases = [dict()]
    if is_float:
        cases += [dict(lo=-1.0, span=2.0), dict(lo=0.0, span=255.0)]
    elif dt != "BOOL":
        top = {"INT8": 200, "UINT8": 256, "INT16": 60000, "UINT16": 65536}.get(dt, 30522)
        cases += [dict(ilo=-5, irange=top), dict(ilo=0, irange=2)]
    for kw in cases:
        for n in (0, 5, 16, 4099):
            a = cref.fill(n, dt, seed=0xABCDEF0123456789, stream=(9 << 32) | 77, **kw)
            b = fill.fill_bytes(n, dt, seed=0xABCDEF0123456789, stream=(9 << 32) | 77, **kw)
            if dt in ("FP32", "FP64") and kw:
                # numpy has no fma: 1 ulp tolerance for scaled float ranges
                m = n - n % 8
                it = np.int32 if dt == "FP32" else np.int64
                d = np.abs(a[:m].view(it).astype(np.int64) - b[:m].view(it).astype(np.int64))
                assert d.size == 0 or d.max() <= 1
            else:
                assert np.array_equal(a, b), (dt, kw, n)


def test_fill_statistics_and_streams():
    v = cref.fill(4 * 200000, "FP32", seed=1, stream=0).view(np.float32)
    assert 0.0 <= v.min() and v.max() < 1.0 and abs(v.mean() - 0.5) < 5e-3
    ids = cref.fill(8 * 50000, "INT64", seed=1, stream=0, ilo=0, irange=30522).view(np.int64)
    assert ids.min() >= 0 and ids.max() < 30522 and len(np.unique(ids)) > 20000
    assert not np.array_equal(cref.fill(4096, "FP32", seed=1, stream=0), cref.fill(4096, "FP32", seed=1, stream=1))
    assert not np.array_equal(cref.fill(4096, "FP32", seed=1, stream=0), cref.fill(4096, "FP32", seed=2, stream=0))
    # a prefix of a longer tensor is the shorter tensor (group g depends on g only)
    assert np.array_equal(cref.fill(10000, "FP16", seed=3, stream=4)[:4096], cref.fill(4096, "FP16", seed=3, stream=4))


def test_codecs_against_reference_goldens(golden):
    for c in golden["codecs"]:
        if c["kind"] == "BYTES":
            items = [bytes.fromhex(x) if not x.startswith("str:") else x[4:] for x in c["items"]]
            if c["dtype"].startswith("|S"):
                arr = np.array(items, dtype=c["dtype"]).reshape(c["shape"])
            else:
                arr = np.empty(len(items), dtype=object)
                for i, it in enumerate(items):
                    arr[i] = int(it) if isinstance(it, str) and it.isdigit() else it
                arr = arr.reshape(c["shape"])
            assert wire.bytes_tensor_wire(arr).hex() == c["wire"]
        else:
            arr = np.frombuffer(bytes.fromhex(c["f32"]), dtype="<f4").reshape(c["shape"])
            assert wire.bf16_tensor_wire(arr).hex() == c["wire"]
    # known-answer vectors quoted from the reference (Rust infer.rs:1095-1106)
    assert wire.bytes_tensor_wire(np.array([b"ab", b"c"], dtype=object)).hex() == "0200000061620100000063"
    hw = wire.bytes_tensor_wire(np.array([b"hello", b"world"], dtype=object))
    assert len(hw) == 18 and hw[:4].hex() == "05000000"
    assert wire.bf16_tensor_wire(np.array([1.0, -2.5], np.float32)).hex() == "803f20c0"


def _build_http_case(name, W):
    a16 = np.arange(16, dtype=np.int32)[None, :]
    m16 = np.full((1, 16), -1, dtype=np.int32)
    if name == "http_config1":
        return [W.HttpInput("INPUT0", [1, 16], "INT32").set_data(a16, False), W.HttpInput("INPUT1", [1, 16], "INT32").set_data(m16)], \
               [W.HttpOutput("OUTPUT0", True), W.HttpOutput("OUTPUT1", False)]
    if name == "http_cudashm_A":
        return [W.HttpInput("INPUT0", [1, 16], "INT32").set_shm("input0_data", 64), W.HttpInput("INPUT1", [1, 16], "INT32").set_shm("input1_data", 64, 64)], \
               [W.HttpOutput("OUTPUT0").set_shm("output0_data", 64), W.HttpOutput("OUTPUT1").set_shm("output1_data", 64)]
    if name in ("http_B_params", "http_seq_string"):
        return [W.HttpInput("INPUT0", [1, 16], "INT32").set_data(a16), W.HttpInput("INPUT1", [1, 16], "INT32").set_data(m16)], None
    if name == "http_C_mixed":
        return [W.HttpInput("S", [1, 2], "BYTES").set_data(np.array([[b"ab", "c"]], dtype=object)),
                W.HttpInput("B", [2], "BF16").set_data(np.array([1.0, -2.5], np.float32)),
                W.HttpInput("H", [2], "FP16").set_data(np.array([1.0, -2.5], np.float16))], [W.HttpOutput("OUT", class_count=3)]
    if name == "http_json_data":
        return [W.HttpInput("B", [2, 2], "BOOL").set_data(np.array([[True, False], [False, True]]), False),
                W.HttpInput("S", [2], "BYTES").set_data(np.array([b"ab", "cd"], dtype=object), False),
                W.HttpInput("U", [3], "UINT64").set_data(np.array([0, 1, 2**64 - 1], dtype=np.uint64), False)], None
    if name == "http_densenet_fp32":
        x = np.random.default_rng(0).random((3, 224, 224), dtype=np.float32)
        return [W.HttpInput("data_0", [3, 224, 224], "FP32").set_data(x)], [W.HttpOutput("fc6_1")]
    raise KeyError(name)


HTTP_CASES = ["http_config1", "http_cudashm_A", "http_B_params", "http_seq_string", "http_C_mixed", "http_json_data", "http_densenet_fp32"]


@pytest.mark.parametrize("name", HTTP_CASES)
def test_http_bodies_against_reference_goldens(golden, name):
    import hashlib

    g = golden[name]
    inputs, outputs = _build_http_case(name, wire)
    body, js = wire.http_request_body(inputs, outputs, **g["kwargs"])
    assert js == g["json_size"]
    assert hashlib.sha256(body).hexdigest() == g["sha256"]
    assert body.hex().startswith(g["body"])


def test_survey_quoted_vectors(golden):
    """The numbers SURVEY.md 9.4 quotes: config 1 = 319/383 B, A = 558 B, B sha256."""
    assert golden["http_config1"]["json_size"] == 319 and len(golden["http_config1"]["body"]) // 2 == 383
    assert golden["http_cudashm_A"]["json_size"] is None and len(golden["http_cudashm_A"]["body"]) // 2 == 558
    assert golden["http_B_params"]["sha256"] == "6bee7338893b8276c9b646bdba32ac40c184dbbf47b098df7666b187c0c19737"
    assert golden["http_C_mixed"]["body"].endswith("0200000061620100000063" + "803f20c0" + "003c00c1")


def test_http_response_parse_against_reference_golden(golden):
    g = golden["http_response"]
    got = wire.http_parse_response(bytes.fromhex(g["body"]), g["header_length"])
    for name, want in g["parsed"].items():
        if want is None:
            assert name not in got
        elif want["dtype"] == "object":
            assert [x.hex() for x in got[name].reshape(-1).tolist()] == want["items"]
        else:
            assert list(got[name].shape) == want["shape"]
            assert np.ascontiguousarray(got[name]).astype(want["dtype"]).tobytes().hex() == want["data"]


def test_grpc_bytes_against_reference_goldens(golden):
    ids = (np.arange(384, dtype=np.int64) * 79 % 30522).reshape(1, 384)
    mask = np.ones((1, 384), dtype=np.int64)
    bert = [wire.GrpcInput("input_ids", [1, 384], "INT64").set_data(ids), wire.GrpcInput("attention_mask", [1, 384], "INT64").set_data(mask)]
    assert wire.grpc_request_bytes("bert_large", bert).hex() == golden["grpc_bert_raw"]["bytes"]
    tok = (np.arange(4096, dtype=np.int32) * 31 % 128256).reshape(1, 4096)
    llama = [wire.GrpcInput("input_ids", [1, 4096], "INT32").set_data(tok)]
    assert wire.grpc_request_bytes("llama3_8b", llama, request_id="42", outputs=["logits"]).hex() == golden["grpc_llama_stream"]["bytes"]
    one = wire.grpc_request_bytes("m", bert[:1], parameters=[("priority", "uint64", 3)])
    assert one.hex() == golden["grpc_one_param"]["bytes"]
    mixed = [wire.GrpcInput("S", [1, 2], "BYTES").set_data(np.array([[b"ab", "c"]], dtype=object)),
             wire.GrpcInput("B", [2], "BF16").set_data(np.array([1.0, -2.5], np.float32)),
             wire.GrpcInput("Z", [0], "FP32").set_data(np.zeros(0, np.float32))]
    assert wire.grpc_request_bytes("m", mixed).hex() == golden["grpc_mixed"]["bytes"]


def test_image_oracle_against_reference_preprocess():
    z = np.load(os.path.join(GOLD, "image_golden.npz"))
    n = 0
    for key in z.files:
        if key.startswith("src_"):
            continue
        dims, dtype, scaling, layout = key.split("_")
        src = z["src_" + dims]
        ref = z[key]
        want = np.frombuffer(np.ascontiguousarray(ref).tobytes(), np.uint8)
        assert ref.dtype == (np.float32 if dtype == "FP32" else np.float16)
        assert np.array_equal(image.pack_batch(src[None], dtype, layout, scaling), want), key
        assert np.array_equal(cref.pack_image(src[None], dtype, layout, scaling), want), key
        n += 1
    assert n == 36


def test_check_oracles():
    a = np.arange(16, dtype=np.int32)
    b = np.ones(16, dtype=np.int32)
    assert cref.addsub_mismatches(a + b, a - b, a, b) == 0
    assert cref.addsub_mismatches(a + b + 1, a - b, a, b) == 16
    s, x = cref.checksum(np.array([1, 2, 3], dtype=np.uint32))
    assert (s, x) == (6, 0)
    s, x = cref.checksum(np.frombuffer(b"\x01\x00\x00\x00\xff", np.uint8))
    assert (s, x) == (1 + 0xFF, 1 ^ 0xFF)
    assert cref.top1(np.array([0.1, np.nan, 3.0, 3.0, -np.inf], np.float32)) == (2, 3.0, 2)


def test_topk_oracle_is_numpy_stable_argsort():
    """oracle_topk == numpy.argsort(-x, kind='stable')[:k]: value descending, lower index
    first on ties (-0.0 == +0.0), NaN after every number, 0xFFFFFFFF past the end."""
    rng = np.random.default_rng(11)
    for _ in range(300):
        n, k = int(rng.integers(1, 80)), int(rng.integers(1, 12))
        x = rng.integers(-3, 4, n).astype(np.float32)
        x[rng.random(n) < 0.15] = np.nan
        x[rng.random(n) < 0.10] = -0.0
        x[rng.random(n) < 0.05] = np.inf
        x[rng.random(n) < 0.05] = -np.inf
        vals, idx = cref.topk(x, k)
        order = np.argsort(-x, kind="stable")[:k]
        m = min(k, n)
        assert np.array_equal(idx[:m], order.astype(np.uint32)[:m])
        assert np.array_equal(vals[:m], x[order[:m]], equal_nan=True)
        assert (idx[m:] == 0xFFFFFFFF).all()


def test_resize_oracle_is_pillow_bilinear():
    """oracle.image.pil_bilinear_resize == Image.resize((w, h), Image.BILINEAR), the call of
    image_client.preprocess (src/python/examples/image_client.py:166): up- and down-scaling,
    one axis only, identity, extreme aspect ratios (pass order switch at 100:1)."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(3)
    cases = [(375, 500, 224, 224), (224, 224, 224, 224), (100, 37, 224, 224), (1080, 1920, 224, 224), (224, 500, 224, 224),
             (300, 224, 224, 224), (17, 23, 5, 7), (5, 7, 17, 23), (600, 600, 299, 299), (1, 1, 4, 4), (449, 449, 224, 224),
             (2, 3, 64, 64), (3000, 30, 64, 48), (3030, 30, 64, 48), (30, 3500, 48, 64)]
    for h, w, oh, ow in cases:
        for c in (3, 1):
            a = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
            pil = np.array(Image.fromarray(a if c == 3 else a[:, :, 0]).resize((ow, oh), Image.BILINEAR))
            if pil.ndim == 2:
                pil = pil[:, :, None]
            assert np.array_equal(image.pil_bilinear_resize(a, oh, ow), pil), (h, w, oh, ow, c)


def _resize_golden_cases():
    z = np.load(os.path.join(GOLD, "image_resize_golden.npz"))
    for key in z.files:
        if key.startswith("src_"):
            continue
        tag, dtype, scaling, layout = key.rsplit("_", 3)
        oh, ow = (int(v) for v in tag.split("_to_")[1].split("x"))
        yield key, z["src_" + tag], z[key], dtype, scaling, layout, oh, ow


def test_resize_and_pack_oracle_against_reference_preprocess():
    """The reference's image_client.preprocess WITH its Image.resize (fixtures generated by
    oracle/gen_golden.py from the reference itself): resize oracle + scaling oracle."""
    n = 0
    for key, src, ref, dtype, scaling, layout, oh, ow in _resize_golden_cases():
        resized = image.pil_bilinear_resize(src, oh, ow)
        want = np.frombuffer(np.ascontiguousarray(ref).tobytes(), np.uint8)
        assert np.array_equal(image.pack_batch(resized[None], dtype, layout, scaling), want), key
        n += 1
    assert n == 12
