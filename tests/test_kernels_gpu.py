"""GPU parity tests: every kernel, through the C ABI, against the CPU oracle.

Integer / byte / index work is compared bit-exactly.  Floating-point random fill
is bit-exact against the C oracle (which uses fmaf like the kernel); against the
numpy oracle the stated tolerance is 1 ulp for scaled ranges (oracle/fill.py).
"""


import numpy as np
import pytest

from oracle import cref
from oracle import fill as ofill
from oracle import image as oimage

pytestmark = pytest.mark.gpu

FLOATS = ("FP32", "FP16", "BF16", "FP64")
ALL_FILL_TYPES = ["FP32", "FP16", "BF16", "FP64", "INT64", "UINT64", "INT32", "UINT32",
                  "INT16", "UINT16", "INT8", "UINT8", "BOOL"]


def _fill_device(ops, nbytes, dt, seed, stream, offset=0, **kw):
    from client_b200.device import DeviceBuffer, make_fill_job

    buf = DeviceBuffer(0, nbytes + offset + 64)
    guard = np.full(nbytes + offset + 64, 0xA5, np.uint8)
    ops.h2d(buf.ptr, guard.ctypes.data, guard.size)
    job = make_fill_job(buf.ptr + offset, nbytes, dt, stream, **kw)
    ops.fill([job], seed=seed)
    out = ops.download(buf.ptr, nbytes + offset + 64)
    assert (out[:offset] == 0xA5).all() and (out[offset + nbytes:] == 0xA5).all(), "fill wrote outside its tensor"
    return out[offset:offset + nbytes].copy()


def _oracle_kwargs(dt, low, high):
    if dt in FLOATS:
        return dict(lo=0.0 if high is None else low, span=0.0 if high is None else high - low)
    if dt == "BOOL" or high is None:
        return {}
    return dict(ilo=low, irange=high - low)


@pytest.mark.parametrize("dt", ALL_FILL_TYPES)
@pytest.mark.parametrize("nbytes", [0, 1, 15, 16, 17, 4096, 16384, 16400, 602112 + 7, 3 * 16384])
def test_fill_default_range_bit_exact(gpu_ops, dt, nbytes):
    es = {"FP64": 8, "INT64": 8, "UINT64": 8, "FP32": 4, "INT32": 4, "UINT32": 4}.get(dt, 2 if dt in ("FP16", "BF16", "INT16", "UINT16") else 1)
    nbytes -= nbytes % es  # whole elements
    got = _fill_device(gpu_ops, nbytes, dt, seed=0x1234567890ABCDEF, stream=(5 << 32) | 17)
    ref = cref.fill(nbytes, dt, seed=0x1234567890ABCDEF, stream=(5 << 32) | 17)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("dt,low,high", [
    ("FP32", -1.0, 1.0), ("FP32", 0.0, 255.0), ("FP16", -1.0, 1.0), ("BF16", -4.0, 4.0), ("FP64", -1.0, 3.0),
    ("INT64", 0, 30522), ("INT64", -7, 1 << 40), ("UINT64", 5, 1 << 63), ("INT32", 0, 128256), ("INT32", -100, 100),
    ("UINT32", 0, 1 << 32), ("INT16", -300, 300), ("UINT16", 0, 65536), ("INT8", -128, 128), ("UINT8", 0, 200),
])
def test_fill_ranges_vs_both_oracles(gpu_ops, dt, low, high):
    nbytes = 100000 - (100000 % 16) + 8
    got = _fill_device(gpu_ops, nbytes, dt, seed=42, stream=7, low=low, high=high)
    kw = _oracle_kwargs(dt, low, high)
    ref_c = cref.fill(nbytes, dt, seed=42, stream=7, **kw)
    assert np.array_equal(got, ref_c), "kernel differs from the C oracle"
    ref_np = ofill.fill_bytes(nbytes, dt, seed=42, stream=7, **kw)
    if dt in ("FP32", "FP64"):
        # stated tolerance vs the numpy oracle (no true fma in numpy): 1 ulp
        vt = np.float32 if dt == "FP32" else np.float64
        it = np.int32 if dt == "FP32" else np.int64
        a = got[: nbytes - nbytes % 8].view(vt).view(it).astype(np.int64)
        b = ref_np[: nbytes - nbytes % 8].view(vt).view(it).astype(np.int64)
        assert np.abs(a - b).max() <= 1
        v = got[: nbytes - nbytes % 8].view(vt)
        assert v.min() >= low and v.max() < high
    else:
        assert np.array_equal(got, ref_np)
    if dt.startswith(("INT", "UINT")):
        npdt = {"INT64": np.int64, "UINT64": np.uint64, "INT32": np.int32, "UINT32": np.uint32,
                "INT16": np.int16, "UINT16": np.uint16, "INT8": np.int8, "UINT8": np.uint8}[dt]
        v = got[: nbytes - nbytes % 8].view(npdt)
        assert int(v.min()) >= low and int(v.max()) < high


def test_fill_unaligned_destination(gpu_ops):
    for off in (1, 3, 4, 8, 12):
        got = _fill_device(gpu_ops, 5000, "UINT8", seed=9, stream=1, offset=off)
        assert np.array_equal(got, cref.fill(5000, "UINT8", seed=9, stream=1))


def test_fill_unaligned_cells_every_offset_and_tail(gpu_ops):
    """fill_unaligned_kernel: every misalignment 1..15 x tensor sizes around the cell
    boundaries (shorter than the first cell, ending inside / exactly at / just after a cell),
    guard bytes on both sides; plus mixed aligned + unaligned jobs in one launch."""
    from client_b200.device import DeviceBuffer, make_fill_job

    for off in range(1, 16):
        for nbytes in (1, 2, 15 - off + 1, 16 - off, 17 - off, 16, 17, 31, 32, 33, 48 - off, 1000, 4096 + off):
            if nbytes <= 0:
                continue
            got = _fill_device(gpu_ops, nbytes, "UINT8", seed=3, stream=off * 131 + nbytes, offset=off)
            assert np.array_equal(got, cref.fill(nbytes, "UINT8", seed=3, stream=off * 131 + nbytes)), (off, nbytes)
    for dt, kw, okw in (("INT64", dict(low=0, high=30522), dict(ilo=0, irange=30522)), ("FP32", {}, {}), ("FP16", dict(low=-1.0, high=1.0), dict(lo=-1.0, span=2.0)),
                        ("INT32", dict(low=0, high=128256), dict(ilo=0, irange=128256)), ("BOOL", {}, {})):
        for off in (3, 5, 10, 13):
            got = _fill_device(gpu_ops, 3072, dt, seed=11, stream=7, offset=off, **kw)
            assert np.array_equal(got, cref.fill(3072, dt, seed=11, stream=7, **okw)), (dt, off)
    got = _fill_device(gpu_ops, 8 * 28, "BYTES", seed=5, stream=2, offset=7, string_length=24)
    assert np.array_equal(got, cref.fill(8 * 28, "BYTES", seed=5, stream=2, irange=24))
    # one launch: aligned job, unaligned jobs (random / zero / byte), sizes that split across CTAs
    size = 300000
    buf = DeviceBuffer(0, 4 * size + 256)
    gpu_ops.h2d(buf.ptr, np.full(4 * size + 256, 0xA5, np.uint8).ctypes.data, 4 * size + 256)
    jobs = [make_fill_job(buf.ptr, size - 16, "FP32", 1), make_fill_job(buf.ptr + size + 3, size - 16, "INT64", 2, low=-5, high=1 << 40),
            make_fill_job(buf.ptr + 2 * size + 9, size - 16, "FP32", 3, mode="zero"), make_fill_job(buf.ptr + 3 * size + 14, size - 16, "FP32", 4, mode="byte", low=0x3C)]
    gpu_ops.fill(jobs, seed=21)
    out = gpu_ops.download(buf.ptr, 4 * size + 256)
    n = size - 16
    assert np.array_equal(out[:n], cref.fill(n, "FP32", seed=21, stream=1))
    assert np.array_equal(out[size + 3:size + 3 + n], cref.fill(n, "INT64", seed=21, stream=2, ilo=-5, irange=(1 << 40) + 5))
    assert (out[2 * size + 9:2 * size + 9 + n] == 0).all() and (out[3 * size + 14:3 * size + 14 + n] == 0x3C).all()
    for a, b in ((n, size + 3), (size + 3 + n, 2 * size + 9), (2 * size + 9 + n, 3 * size + 14), (3 * size + 14 + n, 4 * size + 256)):
        assert (out[a:b] == 0xA5).all(), "fill wrote between its tensors"


def test_fill_zero_and_byte(gpu_ops):
    from client_b200.device import DeviceBuffer, make_fill_job

    buf = DeviceBuffer(0, 70000)
    gpu_ops.fill([make_fill_job(buf.ptr, 70000, "FP32", mode="byte", low=0x5A)])
    assert (gpu_ops.download(buf.ptr, 70000) == 0x5A).all()
    gpu_ops.fill([make_fill_job(buf.ptr + 16, 69000, "FP32", mode="zero")])
    out = gpu_ops.download(buf.ptr, 70000)
    assert (out[:16] == 0x5A).all() and (out[16:69016] == 0).all() and (out[69016:] == 0x5A).all()


def test_fill_many_jobs_one_launch(gpu_ops):
    """One launch, many slots: uniform (64 x 602112 B, config C2) and ragged mixes."""
    from client_b200.device import DeviceBuffer, make_fill_job

    sizes = [602112] * 8 + [3072, 3072, 0, 16384, 5, 100000]
    dts = ["FP32"] * 8 + ["INT64", "INT64", "FP32", "INT32", "UINT8", "FP16"]
    offs = np.concatenate([[0], np.cumsum([(s + 255) // 256 * 256 for s in sizes])])
    buf = DeviceBuffer(0, int(offs[-1]) + 256)
    launches0 = gpu_ops.ctx.launch_count
    jobs = []
    for i, (s, dt) in enumerate(zip(sizes, dts)):
        hi = 30522 if dt == "INT64" else (128256 if dt == "INT32" else None)
        jobs.append(make_fill_job(buf.ptr + int(offs[i]), s, dt, stream_id=1000 + i, low=0 if hi else 0.0, high=hi))
    gpu_ops.fill(jobs, seed=77, epoch=5)
    gpu_ops.sync()
    assert gpu_ops.ctx.launch_count - launches0 == 1
    out = gpu_ops.download(buf.ptr, int(offs[-1]))
    for i, (s, dt) in enumerate(zip(sizes, dts)):
        kw = dict(ilo=0, irange=30522) if dt == "INT64" else (dict(ilo=0, irange=128256) if dt == "INT32" else {})
        ref = cref.fill(s, dt, seed=77, stream=1000 + i + 5, **kw)
        assert np.array_equal(out[int(offs[i]):int(offs[i]) + s], ref), (i, dt, s)
    # uniform fast path
    jobs = [make_fill_job(buf.ptr + i * 602112, 602112, "FP32", stream_id=i) for i in range(8)]
    gpu_ops.fill(jobs, seed=1)
    out = gpu_ops.download(buf.ptr, 8 * 602112)
    for i in range(8):
        assert np.array_equal(out[i * 602112:(i + 1) * 602112], cref.fill(602112, "FP32", seed=1, stream=i))


def _tune(key, value):
    from client_b200 import _native

    _native.check(_native.load().tb200_tune(key, value))


@pytest.mark.parametrize("njobs,nbytes,dt,low,high", [
    (64, 602112, "FP32", None, None),     # C2: interleaved rows, 9 CTAs per tensor, table of 64
    (64, 602112, "FP32", -1.0, 1.0),
    (3, 48, "FP16", None, None),          # one CTA
    (100, 65536, "INT64", 0, 30522),      # balanced ranges across tensors, table of 256
    (256, 3072, "INT32", 0, 128256),      # small tensors, several per CTA
    (37, 16 * 1000, "BF16", -4.0, 4.0),
    (5, 16 * 70001, "UINT8", 0, 200),
    (2, 16 * 300000, "FP64", -1.0, 3.0),
    (7, 16 * 4099, "BOOL", None, None),
    (257, 3072, "INT16", -300, 300),      # more tensors than the parameter table holds: general kernel
])
def test_fill_homogeneous_launch_bit_exact(gpu_ops, njobs, nbytes, dt, low, high):
    """fill_uniform_kernel (dtype-specialised, stream-hoisted Philox, job table in the kernel
    parameters) for every dtype class and both work splits, against the C oracle; and the
    general kernel produces the same bytes for the same launch."""
    from client_b200.device import DeviceBuffer, make_fill_job

    buf = DeviceBuffer(0, njobs * nbytes + 32)
    kw = _oracle_kwargs(dt, low, high)
    rng_kw = {} if high is None else dict(low=low, high=high)
    jobs = [make_fill_job(buf.ptr + 16 + i * nbytes, nbytes, dt, stream_id=(i << 33) | (7 * i + 1), **rng_kw) for i in range(njobs)]
    outs = []
    for uniform in (1, 0):
        _tune(b"fill_uniform", uniform)
        try:
            guard = np.full(njobs * nbytes + 32, 0xA5, np.uint8)
            gpu_ops.h2d(buf.ptr, guard.ctypes.data, guard.size)
            gpu_ops.sync()
            launches0 = gpu_ops.ctx.launch_count
            gpu_ops.fill(jobs, seed=0xC0FFEE1234, epoch=3)
            gpu_ops.sync()
            assert gpu_ops.ctx.launch_count - launches0 == 1
            outs.append(gpu_ops.download(buf.ptr, njobs * nbytes + 32))
        finally:
            _tune(b"fill_uniform", 1)
    out = outs[0]
    assert (out[:16] == 0xA5).all() and (out[-16:] == 0xA5).all(), "fill wrote outside its tensors"
    for i in sorted(set([0, 1, njobs // 2, njobs - 1]) if njobs > 8 else range(njobs)):
        ref = cref.fill(nbytes, dt, seed=0xC0FFEE1234, stream=((i << 33) | (7 * i + 1)) + 3, **kw)
        assert np.array_equal(out[16 + i * nbytes:16 + (i + 1) * nbytes], ref), (i, dt)
    assert np.array_equal(outs[0], outs[1]), "specialised and general kernel disagree"


def test_fill_overlapped_launches_keep_stream_order(gpu_ops):
    """Back-to-back homogeneous fills overlap (programmatic dependent launch) only when they write
    disjoint memory: a rewrite of the same tensors must still win, a rotation over four slot sets
    must leave each set with the data of its last launch, and everything is complete at sync."""
    from client_b200.device import DeviceBuffer, make_fill_job

    slots, n, sets = 16, 602112, 4
    buf = DeviceBuffer(0, sets * slots * n)
    jobs = [[make_fill_job(buf.ptr + (s * slots + k) * n, n, "FP32", stream_id=100 * s + k) for k in range(slots)] for s in range(sets)]
    for trial in range(6):
        gpu_ops.fill(jobs[0], seed=1, epoch=trial)
        gpu_ops.fill(jobs[0], seed=2, epoch=trial)        # same memory: serialised, must win
        gpu_ops.fill(jobs[1], seed=3, epoch=trial)        # disjoint: may overlap
        gpu_ops.fill(jobs[0], seed=4, epoch=trial)        # overlaps the chain's first range again
        gpu_ops.sync()
        a = gpu_ops.download(buf.ptr + (slots - 1) * n, n)
        b = gpu_ops.download(buf.ptr + slots * n, n)
        assert np.array_equal(a, cref.fill(n, "FP32", seed=4, stream=slots - 1 + trial)), trial
        assert np.array_equal(b, cref.fill(n, "FP32", seed=3, stream=100 + trial)), trial
    last = {}
    for i in range(41):
        s = i % sets
        gpu_ops.fill(jobs[s], seed=9, epoch=1000 + i)
        last[s] = 1000 + i
    res = gpu_ops.check_one("sum", buf.ptr + 3 * n, n)  # a plain kernel behind the chain sees finished data
    ref = cref.fill(n, "FP32", seed=9, stream=3 + last[0])
    assert (res["sum"], res["xor32"]) == cref.checksum(ref)
    gpu_ops.sync()
    for s in range(sets):
        for k in (0, slots - 1):
            got = gpu_ops.download(buf.ptr + (s * slots + k) * n, n)
            assert np.array_equal(got, cref.fill(n, "FP32", seed=9, stream=100 * s + k + last[s])), (s, k)


def test_fill_overlap_hazards_are_per_tensor_not_per_hull(gpu_ops):
    """Interleaved slot subsets of ONE region (what consecutive device passes of the load generator
    write) are disjoint byte ranges and may overlap; a subset that shares a slot with a launch
    still in flight may not.  The final bytes are always those of the last writer."""
    from client_b200.device import DeviceBuffer, make_fill_job

    slots, n = 64, 602112
    buf = DeviceBuffer(0, slots * n)
    job = lambda k, sid: make_fill_job(buf.ptr + k * n, n, "FP32", stream_id=sid)  # noqa: E731
    for trial in range(4):
        even = [job(k, 1000 + k) for k in range(0, slots, 2)]
        odd = [job(k, 2000 + k) for k in range(1, slots, 2)]
        third = [job(k, 3000 + k) for k in (0, 5, 63)]          # shares slots with both
        shuffled = [job(k, 4000 + k) for k in (40, 3, 22, 9)]    # unsorted job order, shares slots with `odd`
        gpu_ops.fill(even, seed=6, epoch=trial)
        gpu_ops.fill(odd, seed=6, epoch=trial)
        gpu_ops.fill(third, seed=6, epoch=trial)
        gpu_ops.fill(shuffled, seed=6, epoch=trial)
        gpu_ops.sync()
        owner = {k: (1000 if k % 2 == 0 else 2000) + k for k in range(slots)}
        owner.update({k: 3000 + k for k in (0, 5, 63)})
        owner.update({k: 4000 + k for k in (40, 3, 22, 9)})
        for k in (0, 1, 2, 3, 5, 9, 22, 40, 62, 63):
            got = gpu_ops.download(buf.ptr + k * n, n)
            assert np.array_equal(got, cref.fill(n, "FP32", seed=6, stream=owner[k] + trial)), (trial, k)


def test_graph_chain_of_fills_advances_epoch_per_fill(gpu_ops):
    """Several fill_epoch calls in one capture (overlapping nodes): call i of replay r uses
    epoch0 + (r * calls + i) * bump, exactly the sequence of the same calls issued one by one."""
    from client_b200.device import DeviceBuffer, make_fill_job

    slots, n, sets, calls = 8, 602112, 4, 8
    buf = DeviceBuffer(0, sets * slots * n)
    jobs = [[make_fill_job(buf.ptr + (s * slots + k) * n, n, "FP32", stream_id=k) for k in range(slots)] for s in range(sets)]
    gpu_ops.epoch_set(50)
    gpu_ops.graph_begin()
    for i in range(calls):
        gpu_ops.fill_epoch(jobs[i % sets], seed=11, bump=slots)
    g = gpu_ops.graph_end()
    for r in range(3):
        g.launch()
        gpu_ops.sync()
        for s in range(sets):
            i_last = max(i for i in range(calls) if i % sets == s)
            e = 50 + (r * calls + i_last) * slots
            for k in (0, slots - 1):
                got = gpu_ops.download(buf.ptr + (s * slots + k) * n, n)
                assert np.array_equal(got, cref.fill(n, "FP32", seed=11, stream=k + e)), (r, s, k)
    g.close()


def test_fill_full_size_checksum(gpu_ops):
    """C3 size (38,535,168 B FP16): checksum of the device tensor, computed on the
    device, equals the oracle's checksum of the oracle's tensor."""
    from client_b200.device import DeviceBuffer, make_fill_job

    n = 38535168
    buf = DeviceBuffer(0, n)
    gpu_ops.fill([make_fill_job(buf.ptr, n, "FP16", stream_id=3)], seed=2024)
    res = gpu_ops.check_one("sum", buf.ptr, n)
    ref = cref.fill(n, "FP16", seed=2024, stream=3)
    s, x = cref.checksum(ref)
    assert (res["sum"], res["xor32"]) == (s, x)
    # and a sampled exact comparison of the tail
    tail = gpu_ops.download(buf.ptr + n - 65536, 65536)
    assert np.array_equal(tail, ref[-65536:])


@pytest.mark.parametrize("dt", ["FP16", "FP32", "BF16"])
@pytest.mark.parametrize("scaling", ["NONE", "INCEPTION", "VGG"])
@pytest.mark.parametrize("shape", [(2, 224, 224, 3), (1, 16, 16, 3), (3, 20, 24, 3), (2, 7, 9, 3), (2, 32, 32, 1), (1, 5, 5, 4)])
@pytest.mark.parametrize("layout", ["NCHW", "NHWC"])
def test_pack_image_bit_exact(gpu_ops, dt, scaling, shape, layout):
    from client_b200.device import DeviceBuffer

    if scaling == "VGG" and shape[3] == 4:
        pytest.skip("VGG means are defined for 1 or 3 channels")
    rng = np.random.default_rng(hash((dt, scaling, shape, layout)) & 0xFFFF)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    src = gpu_ops.upload(img)
    es = 4 if dt == "FP32" else 2
    dst = DeviceBuffer(0, img.size * es + 32)
    n, h, w, c = shape
    gpu_ops.pack_image(dst.ptr, dt, layout, src.ptr, n, h, w, c, scaling)
    got = gpu_ops.download(dst.ptr, img.size * es)
    assert np.array_equal(got, cref.pack_image(img, dt, layout, scaling))
    assert np.array_equal(got, oimage.pack_batch(img, dt, layout, scaling))


def test_pack_image_all_256_values(gpu_ops):
    """Every uint8 value through every scaling/dtype: the arithmetic has only 256
    inputs per channel, so this is exhaustive."""
    from client_b200.device import DeviceBuffer

    img = np.arange(256, dtype=np.uint8).reshape(1, 16, 16, 1).repeat(3, axis=3)
    img = np.concatenate([img, img[:, ::-1]], axis=0)
    src = gpu_ops.upload(img)
    for dt in ("FP16", "FP32", "BF16"):
        for scaling in ("NONE", "INCEPTION", "VGG"):
            es = 4 if dt == "FP32" else 2
            dst = DeviceBuffer(0, img.size * es)
            gpu_ops.pack_image(dst.ptr, dt, "NCHW", src.ptr, 2, 16, 16, 3, scaling)
            got = gpu_ops.download(dst.ptr, img.size * es)
            assert np.array_equal(got, oimage.pack_batch(img, dt, "NCHW", scaling)), (dt, scaling)


def test_pack_image_full_size_checksum(gpu_ops):
    """C3(ii): 128 x 224 x 224 x 3 uint8 -> FP16 CHW INCEPTION, compared in full."""
    from client_b200.device import DeviceBuffer

    img = np.random.default_rng(3).integers(0, 256, (128, 224, 224, 3), dtype=np.uint8)
    src = gpu_ops.upload(img)
    dst = DeviceBuffer(0, img.size * 2)
    gpu_ops.pack_image(dst.ptr, "FP16", "NCHW", src.ptr, 128, 224, 224, 3, "INCEPTION")
    res = gpu_ops.check_one("sum", dst.ptr, img.size * 2)
    ref = cref.pack_image(img, "FP16", "NCHW", "INCEPTION")
    assert (res["sum"], res["xor32"]) == cref.checksum(ref)
    assert np.array_equal(gpu_ops.download(dst.ptr, img.size * 2), ref)


CASTS = [("UINT8", "FP32"), ("UINT8", "FP16"), ("UINT8", "BF16"), ("INT8", "FP32"), ("INT16", "FP16"),
         ("UINT16", "FP32"), ("INT32", "FP32"), ("INT32", "FP64"), ("INT32", "INT64"), ("UINT32", "INT64"),
         ("INT64", "INT32"), ("INT64", "FP32"), ("INT64", "FP64"), ("FP16", "FP32"), ("BF16", "FP32"),
         ("FP32", "FP16"), ("FP32", "BF16"), ("FP32", "FP64"), ("FP64", "FP32"), ("FP64", "FP16"),
         ("BOOL", "FP32"), ("FP32", "FP32"), ("INT64", "INT64"), ("UINT8", "UINT8")]
NP = {"BOOL": np.bool_, "UINT8": np.uint8, "INT8": np.int8, "UINT16": np.uint16, "INT16": np.int16,
      "UINT32": np.uint32, "INT32": np.int32, "UINT64": np.uint64, "INT64": np.int64,
      "FP16": np.float16, "FP32": np.float32, "FP64": np.float64}


def _random_values(dt, n, rng):
    if dt == "BF16":
        return (rng.integers(0, 1 << 16, n, dtype=np.uint32).astype(np.uint16))
    if dt == "BOOL":
        return rng.integers(0, 2, n).astype(np.bool_)
    if dt.startswith("FP"):
        bits = {"FP16": 16, "FP32": 32, "FP64": 64}[dt]
        raw = rng.integers(0, 1 << 62, n, dtype=np.uint64).astype({16: np.uint16, 32: np.uint32, 64: np.uint64}[bits])
        v = raw.view(NP[dt])
        special = np.array([0.0, -0.0, 1.0, -2.5, 65504.0, 65520.0, 1e-8, 6e-8, 5.96e-8, np.inf, -np.inf, 3.0e38, 1e-40], dtype=np.float64).astype(NP[dt])
        v = v.copy()
        k = min(n, special.size)
        v[:k] = special[:k]
        v[np.isnan(v)] = 1.5  # NaN payload propagation is not part of the contract
        return v
    info = np.iinfo(NP[dt])
    v = rng.integers(info.min, info.max, n, dtype=np.int64 if info.min < 0 else np.uint64, endpoint=True).astype(NP[dt])
    edge = np.array([info.min, info.max, 0, 1]).astype(NP[dt])
    v[: min(n, 4)] = edge[: min(n, 4)]
    return v


@pytest.mark.parametrize("src_t,dst_t", CASTS)
@pytest.mark.parametrize("n", [0, 1, 7, 4096, 100003])
def test_cast_matches_numpy_astype(gpu_ops, src_t, dst_t, n):
    from client_b200.device import DeviceBuffer

    rng = np.random.default_rng(n + len(src_t) * 13 + len(dst_t))
    v = _random_values(src_t, n, rng)
    if src_t == "BF16":
        as_f32 = (v.astype(np.uint32) << 16).view(np.float32)
        as_f32 = np.where(np.isnan(as_f32), np.float32(1.5), as_f32)
        v = (as_f32.view(np.uint32) >> 16).astype(np.uint16)
        src_vals = (v.astype(np.uint32) << 16).view(np.float32)
    else:
        src_vals = v
    with np.errstate(all="ignore"):
        if dst_t == "BF16":
            ref = (src_vals.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16)
        else:
            ref = src_vals.astype(NP[dst_t])
    src = gpu_ops.upload(v) if n else DeviceBuffer(0, 16)
    dst = DeviceBuffer(0, max(ref.nbytes, 16))
    gpu_ops.cast(dst.ptr, dst_t, src.ptr, src_t, n)
    got = gpu_ops.download(dst.ptr, ref.nbytes)
    assert np.array_equal(got, np.ascontiguousarray(ref).view(np.uint8).reshape(-1)), (src_t, dst_t)


def test_cast_unsupported_pair_fails_loudly(gpu_ops):
    from client_b200 import _native
    from client_b200.device import DeviceBuffer

    a = DeviceBuffer(0, 64)
    with pytest.raises(_native.NativeError):
        gpu_ops.cast(a.ptr, "UINT8", a.ptr, "FP32", 4)


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.int64])
def test_pack_strided_equals_tobytes(gpu_ops, dtype):
    from client_b200.device import DeviceBuffer

    rng = np.random.default_rng(5)
    base = rng.integers(0, 100, (6, 10, 14, 9)).astype(dtype)
    src = gpu_ops.upload(base)
    views = [base.transpose(0, 3, 1, 2), base[:, ::2, 1:13:3, ::-1], base[2:5, :, :, 4], base.T,
             base[::-1, ::-1], np.broadcast_to(base[:1, :1], (3, 4, 14, 9)), base[1, 2, 3, 4:5], base[:, :0]]
    for v in views:
        nbytes = v.size * v.itemsize
        dst = DeviceBuffer(0, nbytes + 32)
        off = v.__array_interface__["data"][0] - base.__array_interface__["data"][0]
        gpu_ops.pack_strided(dst.ptr, src.ptr + off, v.itemsize, v.shape, v.strides)
        assert np.array_equal(gpu_ops.download(dst.ptr, nbytes), np.frombuffer(v.tobytes(), np.uint8))


def test_concat_equals_bytes_join(gpu_ops):
    from client_b200._native import CopyJob
    from client_b200.device import DeviceBuffer

    rng = np.random.default_rng(11)
    parts = [rng.integers(0, 256, n, dtype=np.uint8) for n in (319, 64, 0, 16384, 100001, 7, 602112)]
    srcs = [gpu_ops.upload(p) if p.size else DeviceBuffer(0, 16) for p in parts]
    total = sum(p.size for p in parts)
    dst = DeviceBuffer(0, total + 16)
    jobs, off = [], 0
    for p, s in zip(parts, srcs):
        jobs.append(CopyJob(dst=dst.ptr + off, src=s.ptr, nbytes=p.size))
        off += p.size
    gpu_ops.concat(jobs)
    assert gpu_ops.download(dst.ptr, total).tobytes() == b"".join(p.tobytes() for p in parts)


def test_check_kinds(gpu_ops):
    from client_b200._native import CheckJob
    from client_b200.device import HostBuffer, results_array

    rng = np.random.default_rng(2)
    a = rng.integers(-1000, 1000, (1, 16), dtype=np.int32)
    b = rng.integers(-1000, 1000, (1, 16), dtype=np.int32)
    in0, in1 = gpu_ops.upload(a), gpu_ops.upload(b)
    good0, good1 = gpu_ops.upload(a + b), gpu_ops.upload(a - b)
    bad = (a + b).copy()
    bad[0, 3] += 1
    bad0 = gpu_ops.upload(bad)
    logits = rng.standard_normal(1000).astype(np.float32)
    logits[17] = np.nan
    logits[500] = np.inf
    logits[600] = -np.inf
    lg = gpu_ops.upload(logits)
    blob = rng.integers(0, 256, 3 * (1 << 20) + 13, dtype=np.uint8)
    blob2 = blob.copy()
    blob2[[5, 1 << 20, blob.size - 1]] ^= 0xFF
    d1, d2 = gpu_ops.upload(blob), gpu_ops.upload(blob2)
    jobs = [
        CheckJob(a=good0.ptr, b=good1.ptr, c=in0.ptr, d=in1.ptr, nbytes=64, kind=2),
        CheckJob(a=bad0.ptr, b=good1.ptr, c=in0.ptr, d=in1.ptr, nbytes=64, kind=2),
        CheckJob(a=lg.ptr, nbytes=4000, kind=3),
        CheckJob(a=d1.ptr, b=d2.ptr, nbytes=blob.size, kind=1),
        CheckJob(a=d1.ptr + 1, nbytes=blob.size - 1, kind=0),  # unaligned checksum
        CheckJob(a=d1.ptr, nbytes=0, kind=0),
    ]
    res = HostBuffer(4096)
    gpu_ops.check(jobs, res.device_ptr)
    gpu_ops.sync()
    r = results_array(res, len(jobs))
    assert r[0]["mismatches"] == 0 and r[1]["mismatches"] == 1
    assert cref.addsub_mismatches(bad, a - b, a, b) == 1
    assert (int(r[0]["sum"]), int(r[0]["xor32"])) == cref.checksum(a + b)
    idx, mv, nonfinite = cref.top1(logits)
    assert int(r[2]["argmax"]) == idx == 500 and int(r[2]["mismatches"]) == nonfinite == 3
    assert np.isinf(r[2]["max_value"])
    assert int(r[3]["mismatches"]) == cref.count_diff_bytes(blob, blob2) == 3
    assert (int(r[3]["sum"]), int(r[3]["xor32"])) == cref.checksum(blob)
    assert (int(r[4]["sum"]), int(r[4]["xor32"])) == cref.checksum(blob[1:])
    assert int(r[5]["sum"]) == 0 and int(r[5]["mismatches"]) == 0
    # finite logits: argmax equals numpy's
    logits2 = rng.standard_normal(1000).astype(np.float32)
    dlog = gpu_ops.upload(logits2)  # keep the buffer alive for the launch
    out = gpu_ops.check_one("top1", dlog.ptr, 4000)
    assert out["argmax"] == int(np.argmax(logits2)) and out["max_value"] == float(logits2.max())


def test_graph_replay_advances_epoch(gpu_ops):
    """A captured fill produces fresh data on every replay (device epoch), and
    equals the oracle for stream = slot + epoch."""
    from client_b200.device import DeviceBuffer, make_fill_job

    n = 65536 + 48
    buf = DeviceBuffer(0, 2 * n)
    jobs = [make_fill_job(buf.ptr, n, "FP32", stream_id=10), make_fill_job(buf.ptr + n, n, "INT32", stream_id=11, low=0, high=1000)]
    gpu_ops.epoch_set(100)
    gpu_ops.graph_begin()
    gpu_ops.fill_epoch(jobs, seed=5)
    gpu_ops.epoch_bump(2)
    g = gpu_ops.graph_end()
    for it in range(3):
        g.launch()
        out = gpu_ops.download(buf.ptr, 2 * n)
        e = 100 + 2 * it
        assert np.array_equal(out[:n], cref.fill(n, "FP32", seed=5, stream=10 + e))
        assert np.array_equal(out[n:], cref.fill(n, "INT32", seed=5, stream=11 + e, ilo=0, irange=1000))
    g.close()


def test_graph_with_folded_bump_and_parallel_validate(gpu_ops):
    """The bench step: fill (advancing the epoch inside the kernel) on the main branch,
    validation of the output regions on a parallel branch of the same graph."""
    from client_b200._native import CheckJob
    from client_b200.device import DeviceBuffer, HostBuffer, make_fill_job, results_array

    slots, n = 6, 602112
    buf = DeviceBuffer(0, slots * n)
    logits = np.random.default_rng(4).standard_normal((slots, 1000)).astype(np.float32)
    outs = gpu_ops.upload(logits)
    res = HostBuffer(slots * 32)
    jobs = [make_fill_job(buf.ptr + k * n, n, "FP32", stream_id=k) for k in range(slots)]
    checks = [CheckJob(a=outs.ptr + k * 4000, nbytes=4000, kind=3) for k in range(slots)]
    gpu_ops.epoch_set(7)
    launches0 = gpu_ops.ctx.launch_count
    gpu_ops.graph_begin()
    gpu_ops.fork()
    gpu_ops.check(checks, res.device_ptr)           # side branch
    gpu_ops.select(False)
    gpu_ops.fill_epoch(jobs, seed=99, bump=slots)   # main branch, concurrent with the check
    gpu_ops.join()
    g = gpu_ops.graph_end()
    assert gpu_ops.ctx.launch_count == launches0  # captured, not run
    for it in range(4):
        g.launch()
        gpu_ops.sync()
        e = 7 + slots * it
        got = gpu_ops.download(buf.ptr, slots * n)
        for k in (0, slots - 1):
            assert np.array_equal(got[k * n:(k + 1) * n], cref.fill(n, "FP32", seed=99, stream=k + e)), (it, k)
        r = results_array(res, slots)
        assert [int(x) for x in r["argmax"]] == [int(np.argmax(row)) for row in logits]
    assert gpu_ops.ctx.launch_count - launches0 == 4 * 3  # fill, check and the node that advances the device epoch
    g.close()
    # eager fork/join as well
    gpu_ops.fork()
    gpu_ops.check(checks, res.device_ptr)
    gpu_ops.join()
    gpu_ops.fill(jobs, seed=1)
    gpu_ops.sync()
    assert np.array_equal(gpu_ops.download(buf.ptr, n), cref.fill(n, "FP32", seed=1, stream=0))


def test_fill_into_pinned_host_wire_buffer(gpu_ops):
    """C4 shape: 2 x INT64[1,384] per request written by the kernel straight into
    mapped host memory (the gRPC raw_input_contents staging), 256 slots, one launch."""
    from client_b200.device import HostBuffer, make_fill_job

    slots, per = 256, 3072
    hb = HostBuffer(slots * 2 * per)
    jobs = []
    for s in range(slots):
        jobs.append(make_fill_job(hb.device_ptr + (2 * s) * per, per, "INT64", stream_id=2 * s, low=0, high=30522))
        jobs.append(make_fill_job(hb.device_ptr + (2 * s + 1) * per, per, "INT64", stream_id=2 * s + 1, low=0, high=2))
    gpu_ops.fill(jobs, seed=3)
    gpu_ops.sync()
    host = hb.array(np.uint8)
    for s in (0, 1, 100, 255):
        assert np.array_equal(host[(2 * s) * per:(2 * s + 1) * per], cref.fill(per, "INT64", seed=3, stream=2 * s, ilo=0, irange=30522))
        assert np.array_equal(host[(2 * s + 1) * per:(2 * s + 2) * per], cref.fill(per, "INT64", seed=3, stream=2 * s + 1, ilo=0, irange=2))
    ids = host[:per].view(np.int64)
    assert ids.min() >= 0 and ids.max() < 30522


def test_step_sync_call(gpu_ops):
    """tb200_step_sync: one call = generate inputs || validate outputs, both complete on return."""
    from client_b200._native import CheckJob
    from client_b200.device import DeviceBuffer, HostBuffer, make_fill_job, results_array

    n = 602112
    buf = DeviceBuffer(0, 2 * n)
    logits = np.random.default_rng(8).standard_normal((2, 1000)).astype(np.float32)
    outs = gpu_ops.upload(logits)
    res = HostBuffer(64)
    jobs = [make_fill_job(buf.ptr + k * n, n, "FP32", stream_id=k) for k in range(2)]
    checks = [CheckJob(a=outs.ptr + k * 4000, nbytes=4000, kind=3) for k in range(2)]
    gpu_ops.step(jobs, checks, res.device_ptr, seed=3, epoch=9)
    r = results_array(res, 2)
    assert [int(x) for x in r["argmax"]] == [int(np.argmax(row)) for row in logits]
    got = gpu_ops.download(buf.ptr, 2 * n)
    assert np.array_equal(got[n:], cref.fill(n, "FP32", seed=3, stream=1 + 9))
    gpu_ops.step(jobs, [], res.device_ptr, seed=4)  # no validation this time
    assert np.array_equal(gpu_ops.download(buf.ptr, n), cref.fill(n, "FP32", seed=4, stream=0))


def test_step_submit_wait_pipeline(gpu_ops):
    """tb200_step_submit / tb200_step_wait: many steps in flight (more than the ring holds, so
    submits displace and wait for old tickets; more uploads than the job-table ring holds, so slots
    are reclaimed through waited steps), every step's tensors and results exact."""
    from client_b200._native import CheckJob
    from client_b200.device import DeviceBuffer, HostBuffer, make_fill_job, results_array

    n, sets = 48000, 12
    buf = DeviceBuffer(0, sets * n)
    logits = np.random.default_rng(9).standard_normal((sets, 1000)).astype(np.float32)
    outs = gpu_ops.upload(logits)
    res = HostBuffer(sets * 32)
    tickets = []
    for i in range(100):
        s = i % sets
        if len(tickets) >= 10:  # > TB200_STEP_DEPTH: the oldest were displaced (and waited for) by submits
            t, s0, i0 = tickets.pop(0)
            gpu_ops.step_wait(t)
            assert int(results_array(res, sets)["argmax"][s0]) == int(np.argmax(logits[s0]))
            if i0 % 17 == 0:
                assert np.array_equal(gpu_ops.download(buf.ptr + s0 * n, n), cref.fill(n, "FP32", seed=5, stream=s0 + i0))
        t = gpu_ops.step_submit([make_fill_job(buf.ptr + s * n, n, "FP32", stream_id=s)],
                                [CheckJob(a=outs.ptr + s * 4000, nbytes=4000, kind=3)], res.device_ptr + s * 32, seed=5, epoch=i)
        tickets.append((t, s, i))
    for t, s0, i0 in tickets:
        gpu_ops.step_wait(t)
        gpu_ops.step_wait(t)  # idempotent
        assert np.array_equal(gpu_ops.download(buf.ptr + s0 * n, n), cref.fill(n, "FP32", seed=5, stream=s0 + i0))
    with pytest.raises(Exception):
        gpu_ops.step_wait(10**9)
    t = gpu_ops.step_submit([make_fill_job(buf.ptr, n, "FP32", stream_id=1)], [], res.device_ptr, seed=6)  # no validation
    gpu_ops.step_wait(t)
    assert np.array_equal(gpu_ops.download(buf.ptr, n), cref.fill(n, "FP32", seed=6, stream=1))


# ---- resize + pack (image_client.preprocess with its Image.resize) ---------------------------
def _resize_on_device(ops, src, dtype, layout, scaling, oh, ow):
    from client_b200.device import DeviceBuffer

    n, sh, sw, c = src.shape
    es = {"FP32": 4, "FP16": 2, "BF16": 2, "UINT8": 1}[dtype]
    dsrc = ops.upload(src)
    dst = DeviceBuffer(0, n * oh * ow * c * es)
    ops.resize_pack_image(dst.ptr, dtype, layout, dsrc.ptr, n, sh, sw, c, oh, ow, scaling)
    ops.sync()
    return ops.download(dst.ptr, n * oh * ow * c * es)


def test_resize_pack_against_reference_goldens():
    """Device resize + cast + scaling + layout == the reference's preprocess output
    (tests/golden/image_resize_golden.npz, generated from image_client.preprocess)."""
    from client_b200 import _native
    from client_b200.device import DeviceOps
    from test_oracle import _resize_golden_cases

    ops = DeviceOps(_native.default_context(0))
    n = 0
    for key, src, ref, dtype, scaling, layout, oh, ow in _resize_golden_cases():
        got = _resize_on_device(ops, src[None], dtype, layout, scaling, oh, ow)
        assert np.array_equal(got, np.frombuffer(np.ascontiguousarray(ref).tobytes(), np.uint8)), key
        n += 1
    assert n == 12


@pytest.mark.parametrize("shape", [(375, 500, 224, 224), (224, 224, 224, 224), (100, 37, 224, 224), (1080, 1920, 224, 224),
                                   (224, 500, 224, 224), (300, 224, 224, 224), (17, 23, 5, 7), (5, 7, 17, 23), (600, 600, 299, 299),
                                   (1, 1, 4, 4), (449, 449, 224, 224), (2900, 30, 61, 48), (30, 3500, 48, 64), (4000, 3000, 224, 224)])
def test_resize_bare_u8_matches_pillow_pinned_oracle(shape):
    """UINT8 output = the bare Image.resize((w, h), BILINEAR): random pixels, batches of 2,
    c = 1 and 3, up/down/one-axis/identity/extreme scales."""
    from client_b200 import _native
    from client_b200.device import DeviceOps
    from oracle import image

    sh, sw, oh, ow = shape
    ops = DeviceOps(_native.default_context(0))
    rng = np.random.default_rng(sh * 7 + sw)
    for c in (3, 1):
        src = rng.integers(0, 256, (2, sh, sw, c), dtype=np.uint8)
        got = _resize_on_device(ops, src, "UINT8", "NHWC", "NONE", oh, ow).reshape(2, oh, ow, c)
        for i in range(2):
            assert np.array_equal(got[i], image.pil_bilinear_resize(src[i], oh, ow)), (shape, c, i)


def test_resize_pack_dtypes_layouts_and_errors():
    from client_b200 import _native
    from client_b200.device import DeviceBuffer, DeviceOps
    from oracle import cref, image

    ops = DeviceOps(_native.default_context(0))
    rng = np.random.default_rng(9)
    src = rng.integers(0, 256, (3, 90, 130, 3), dtype=np.uint8)
    resized = np.stack([image.pil_bilinear_resize(s, 64, 48) for s in src])
    for dtype in ("FP32", "FP16", "BF16"):
        for layout in ("NCHW", "NHWC"):
            for scaling in ("NONE", "INCEPTION", "VGG"):
                got = _resize_on_device(ops, src, dtype, layout, scaling, 64, 48)
                assert np.array_equal(got, cref.pack_image(resized, dtype, layout, scaling)), (dtype, layout, scaling)
    d = DeviceBuffer(0, 1 << 20)
    with pytest.raises(_native.NativeError, match="100:1"):
        ops.resize_pack_image(d.ptr, "FP32", "NCHW", d.ptr, 1, 3100, 30, 3, 8, 8, "NONE")
    with pytest.raises(_native.NativeError, match="no scaling"):
        ops.resize_pack_image(d.ptr, "UINT8", "NCHW", d.ptr, 1, 30, 30, 3, 8, 8, "VGG")
    with pytest.raises(_native.NativeError, match="bad shape"):
        ops.resize_pack_image(d.ptr, "FP32", "NCHW", d.ptr, 1, 30, 30, 2, 8, 8, "NONE")
    with pytest.raises(_native.NativeError, match="shared memory"):
        ops.resize_pack_image(d.ptr, "FP32", "NCHW", d.ptr, 1, 100000, 1001, 3, 8, 8, "NONE")


def test_set_shared_memory_region_from_image_with_resize():
    import client_b200.utils.cuda_shared_memory as cudashm
    from oracle import cref, image

    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (375, 500, 3), dtype=np.uint8)
    h = cudashm.create_shared_memory_region("resize_in", 3 * 224 * 224 * 4, 0)
    cudashm.set_shared_memory_region_from_image(h, img, "FP32", "INCEPTION", resize=(224, 224))
    got = cudashm.get_contents_as_numpy(h, np.float32, [3, 224, 224])
    want = cref.pack_image(image.pil_bilinear_resize(img, 224, 224)[None], "FP32", "NCHW", "INCEPTION")
    assert np.array_equal(got.view(np.uint8).reshape(-1), want)
    with pytest.raises(cudashm.CudaSharedMemoryException):
        cudashm.set_shared_memory_region_from_image(h, img, "FP32", "INCEPTION", resize=(448, 448))
    cudashm.destroy_shared_memory_region(h)


# ---- BYTES tensors of fixed-length strings generated in place ------------------------------
@pytest.mark.parametrize("count,length", [(1, 0), (5, 1), (7, 13), (1000, 128), (3, 1000), (64, 12)])
def test_fill_bytes_strings(gpu_ops, count, length):
    from client_b200.device import DeviceBuffer, make_fill_job
    from client_b200.utils import deserialize_bytes_tensor

    n = count * (4 + length)
    buf = DeviceBuffer(0, n + 64)
    jobs = [make_fill_job(buf.ptr, n, "BYTES", stream_id=9, string_length=length),
            make_fill_job(buf.ptr + ((n + 15) // 16) * 16, 16, "INT32", stream_id=1)]  # mixed launch
    gpu_ops.fill(jobs, seed=21)
    gpu_ops.sync()
    got = gpu_ops.download(buf.ptr, n)
    assert np.array_equal(got, cref.fill(n, "BYTES", seed=21, stream=9, irange=length))
    strings = deserialize_bytes_tensor(got.tobytes())
    assert len(strings) == count and all(len(s) == length for s in strings)


def test_fill_bytes_rejects_bad_sizes(gpu_ops):
    from client_b200 import _native
    from client_b200._native import FillJob
    from client_b200.device import DeviceBuffer

    buf = DeviceBuffer(0, 4096)
    bad = FillJob(dst=buf.ptr, nbytes=100, stream=0, dtype=_native.DTYPE_CODES["BYTES"], mode=0, irange=13)
    with pytest.raises(_native.NativeError, match="BYTES fill needs"):
        gpu_ops.fill([bad], seed=1)
    zero = FillJob(dst=buf.ptr, nbytes=68, stream=0, dtype=_native.DTYPE_CODES["BYTES"], mode=1, irange=13)
    with pytest.raises(_native.NativeError, match="TB200_FILL_RANDOM only"):
        gpu_ops.fill([zero], seed=1)


# ---- device deflate: decoded by the reference's decompressors ------------------------------
@pytest.mark.parametrize("algorithm", ["gzip", "deflate"])
def test_device_deflate_round_trip(gpu_ops, algorithm):
    import gzip
    import zlib

    from test_host_emul import _deflate_inputs

    for label, data in _deflate_inputs():
        src = gpu_ops.upload(np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8))
        stream = gpu_ops.deflate(src.ptr, len(data), algorithm)
        back = gzip.decompress(stream) if algorithm == "gzip" else zlib.decompress(stream)
        assert back == data, label


def test_device_deflate_of_generated_tensors(gpu_ops):
    """What the wire path would send: generated tensors compressed where they were generated.
    Large input (4704 chunks), zero data, token ids, random floats (stored fallback)."""
    import zlib

    from client_b200.device import DeviceBuffer, make_fill_job

    n = 38535168
    buf = DeviceBuffer(0, n)
    cases = [("zero", make_fill_job(buf.ptr, n, "FP32", mode="zero"), 0.04),
             ("ids", make_fill_job(buf.ptr, n, "INT64", stream_id=3, low=0, high=30522), 0.55),
             ("fp32", make_fill_job(buf.ptr, n, "FP32", stream_id=4), 1.01)]
    for label, job, max_ratio in cases:
        gpu_ops.fill([job], seed=8)
        gpu_ops.sync()
        stream = gpu_ops.deflate(buf.ptr, n, "deflate")
        raw = gpu_ops.download(buf.ptr, n).tobytes()
        assert zlib.decompress(stream) == raw, label
        assert len(stream) <= n * max_ratio, (label, len(stream))


def test_device_deflate_rejects_small_destination(gpu_ops):
    from client_b200 import _native
    from client_b200.device import DeviceBuffer, HostBuffer

    src = DeviceBuffer(0, 1 << 20)
    dst = DeviceBuffer(0, 1 << 20)
    size = HostBuffer(64)
    with pytest.raises(_native.NativeError, match="tb200_deflate_bound"):
        gpu_ops.deflate_async(dst.ptr, 1 << 20, src.ptr, 1 << 20, size.device_ptr, "gzip")


def _serialized(strings):
    from client_b200.utils import serialize_byte_tensor

    arr = np.array(strings, dtype=object)
    ser = serialize_byte_tensor(arr)
    return np.frombuffer(ser.item() if ser.size else b"", dtype=np.uint8)


@pytest.mark.parametrize("case", ["digits", "empty_and_long", "many", "window_edges", "single"])
def test_bytes_decode_on_device_matches_deserialize_bytes_tensor(gpu_ops, case):
    """tb200_bytes_decode_async against utils.deserialize_bytes_tensor (reference
    PY/utils/__init__.py:264-291) on the same serialised bytes: element boundaries (offsets) and
    payloads, for strings shorter / longer than the 32 KiB staging window, empty strings, a header
    straddling a window edge, an unaligned source, and trailing garbage after the last element."""
    from client_b200.device import DeviceBuffer
    from client_b200.utils import deserialize_bytes_tensor

    rng = np.random.default_rng(11)
    if case == "digits":
        strings = [str(i).encode() for i in range(16)]  # the reference's own cudashm BYTES test values
    elif case == "empty_and_long":
        strings = [b"", b"a", b"", rng.bytes(100000), b"tail", rng.bytes(32768), b""]
    elif case == "many":
        strings = [rng.bytes(int(n)) for n in rng.integers(0, 40, 50000)]
    elif case == "window_edges":
        strings = [rng.bytes(32768 - 4 - 2), b"xy", rng.bytes(32768 - 7), b"", rng.bytes(5)]  # headers land across window ends
    else:
        strings = [b"only"]
    ser = _serialized(strings)
    for lead in (0, 3):  # source 16-byte aligned / unaligned
        buf = DeviceBuffer(0, ser.size + 64)
        host = np.concatenate([np.full(lead, 0xEE, np.uint8), ser, np.full(61 - lead, 0xEE, np.uint8)])
        gpu_ops.h2d(buf.ptr, host.ctypes.data, host.size)
        gpu_ops.sync()
        offsets, packed, consumed = gpu_ops.bytes_decode(buf.ptr + lead, ser.size + 20, len(strings))
        want = deserialize_bytes_tensor(ser.tobytes())
        got = [packed[offsets[i]:offsets[i + 1]] for i in range(len(strings))]
        assert consumed == ser.size and len(packed) == sum(len(x) for x in strings)
        assert got == list(want) == strings


def test_bytes_decode_reports_truncated_and_inconsistent_streams(gpu_ops):
    from client_b200.device import DeviceBuffer

    ser = _serialized([b"abc", b"defgh"])
    buf = DeviceBuffer(0, 64)
    gpu_ops.h2d(buf.ptr, ser.ctypes.data, ser.size)
    gpu_ops.sync()
    with pytest.raises(ValueError, match="truncated"):
        gpu_ops.bytes_decode(buf.ptr, ser.size, 3)           # a third element is asked for
    with pytest.raises(ValueError, match="inconsistent"):
        gpu_ops.bytes_decode(buf.ptr, ser.size - 2, 2)       # the second payload runs past the buffer
    offsets, packed, consumed = gpu_ops.bytes_decode(buf.ptr, ser.size, 0)
    assert list(offsets) == [0] and packed == b"" and consumed == 0


def test_get_contents_as_numpy_bytes_goes_through_the_device_decode(gpu_ops):
    """The drop-in's BYTES read-back (reference cuda_shared_memory/__init__.py:306-323): same
    object array as the reference flow, for the values of the reference's own unit test and for a
    2-D shape."""
    import client_b200.utils.cuda_shared_memory as cudashm
    from client_b200.utils import serialize_byte_tensor

    arr = np.array([str(i).encode() for i in range(16)], dtype=object).reshape(4, 4)
    ser = serialize_byte_tensor(arr)
    h = cudashm.create_shared_memory_region("bytes_decode_region", ser.item().__len__() + 32, 0)
    try:
        launches0 = gpu_ops.ctx.launch_count
        cudashm.set_shared_memory_region(h, [ser])
        got = cudashm.get_contents_as_numpy(h, np.object_, [4, 4])
        assert got.dtype == np.object_ and got.shape == (4, 4) and (got == arr).all()
        assert gpu_ops.ctx.launch_count - launches0 >= 2  # scan + gather ran on the device
    finally:
        cudashm.destroy_shared_memory_region(h)
