"""The reference's own cuda shared memory unit tests, restated against the
drop-in module (reference: src/python/library/tests/test_cuda_shared_memory.py:42-164),
plus the error contract and the device-side producers."""

import base64
import multiprocessing as mp

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def cudashm():
    import client_b200.utils.cuda_shared_memory as m

    return m


def test_dlpack_from_gpu(cudashm):
    torch = pytest.importorskip("torch")
    gpu_tensor = torch.ones(4, 4).cuda(0)
    h = cudashm.create_shared_memory_region("cudashm_data", 64, 0)
    cudashm.set_shared_memory_region_from_dlpack(h, [gpu_tensor])
    smt = cudashm.as_shared_memory_tensor(h, "FP32", [4, 4])
    generated = torch.from_dlpack(smt)
    assert torch.allclose(gpu_tensor, generated)
    del generated
    cudashm.destroy_shared_memory_region(h)


def test_dlpack_from_cpu(cudashm):
    torch = pytest.importorskip("torch")
    cpu_tensor = np.ones([4, 4], dtype=np.float32)
    h = cudashm.create_shared_memory_region("cudashm_data", 64, 0)
    cudashm.set_shared_memory_region_from_dlpack(h, [cpu_tensor])
    smt = cudashm.as_shared_memory_tensor(h, "FP32", [4, 4])
    generated = torch.from_dlpack(smt)
    assert np.allclose(cpu_tensor, np.from_dlpack(generated.cpu()))
    del generated
    cudashm.destroy_shared_memory_region(h)


def test_numpy_set_then_dlpack_read(cudashm):
    torch = pytest.importorskip("torch")
    cpu_tensor = np.arange(16, dtype=np.float32).reshape(4, 4)
    h = cudashm.create_shared_memory_region("cudashm_data", 64, 0)
    cudashm.set_shared_memory_region(h, [cpu_tensor])
    generated = torch.from_dlpack(cudashm.as_shared_memory_tensor(h, "FP32", [4, 4]))
    assert np.array_equal(cpu_tensor, generated.cpu().numpy())
    del generated
    cudashm.destroy_shared_memory_region(h)


def test_numpy_round_trip(cudashm):
    cpu_tensor = np.ones([4, 4], dtype=np.float32)
    h = cudashm.create_shared_memory_region("cudashm_data", 64, 0)
    cudashm.set_shared_memory_region(h, [cpu_tensor])
    out = cudashm.get_contents_as_numpy(h, np.float32, [4, 4])
    assert np.allclose(cpu_tensor, out)
    cudashm.destroy_shared_memory_region(h)


def test_numpy_bytes(cudashm):
    import client_b200.utils as utils

    int_tensor = np.arange(start=0, stop=16, dtype=np.int32)
    bytes_tensor = np.array([str(x).encode("utf-8") for x in int_tensor.flatten()], dtype=object)
    bytes_tensor = bytes_tensor.reshape(int_tensor.shape)
    serialized = utils.serialize_byte_tensor(bytes_tensor)
    byte_size = utils.serialized_byte_size(serialized)
    h = cudashm.create_shared_memory_region("cudashm_data", byte_size, 0)
    cudashm.set_shared_memory_region(h, [serialized])
    out = cudashm.get_contents_as_numpy(h, np.object_, [16])
    assert np.array_equal(bytes_tensor, out)
    cudashm.destroy_shared_memory_region(h)


def test_multiple_arrays_offsets_and_large_copy(cudashm):
    """simple_http_cudashm_client.py layout: two int32[1,16] back to back; and a
    38.5 MB tensor (C3) through the chunked pinned staging."""
    a = np.arange(16, dtype=np.int32).reshape(1, 16)
    b = np.ones((1, 16), dtype=np.int32)
    h = cudashm.create_shared_memory_region("input_data", 128, 0)
    cudashm.set_shared_memory_region(h, [a, b])
    out = cudashm.get_contents_as_numpy(h, np.int32, [2, 16])
    assert np.array_equal(out[0], a[0]) and np.array_equal(out[1], b[0])
    cudashm.destroy_shared_memory_region(h)

    big = np.random.default_rng(0).integers(0, 1 << 16, (128, 3, 224, 224), dtype=np.uint16).view(np.float16)
    h = cudashm.create_shared_memory_region("big", big.nbytes, 0)
    cudashm.set_shared_memory_region(h, [big[:64], big[64:]])
    out = cudashm.get_contents_as_numpy(h, np.float16, big.shape)
    assert np.array_equal(out.view(np.uint16), big.view(np.uint16))
    # non-contiguous input is packed in C order like np.ascontiguousarray().flatten()
    t = big[0].transpose(1, 2, 0)
    cudashm.set_shared_memory_region(h, [t])
    out = cudashm.get_contents_as_numpy(h, np.float16, t.shape)
    assert np.array_equal(out.view(np.uint16), np.ascontiguousarray(t).view(np.uint16))
    cudashm.destroy_shared_memory_region(h)


def test_error_contract(cudashm):
    E = cudashm.CudaSharedMemoryException
    h = cudashm.create_shared_memory_region("r", 64, 0)
    with pytest.raises(E, match="input_values must be specified as a numpy array"):
        cudashm.set_shared_memory_region(h, np.zeros(4))
    with pytest.raises(E, match="list/tuple of numpy arrays"):
        cudashm.set_shared_memory_region(h, [[1, 2]])
    with pytest.raises(E, match="unable to set values in cuda shared memory"):
        cudashm.set_shared_memory_region(h, [np.zeros(17, np.float32)])
    with pytest.raises(E, match="insufficient to provide numpy array"):
        cudashm.get_contents_as_numpy(h, np.float32, [17])
    with pytest.raises(E, match="unable to create cuda shared memory handle"):
        cudashm.create_shared_memory_region("bad", 64, 4096)
    assert h in cudashm.allocated_shared_memory_regions()
    raw = cudashm.get_raw_handle(h)
    assert len(base64.b64decode(raw)) == 64
    assert (h._triton_shm_name, h._byte_size, h._device_id) == ("r", 64, 0) and h._base_addr != 0
    cudashm.destroy_shared_memory_region(h)
    assert h not in cudashm.allocated_shared_memory_regions()


def _peer_reads_and_writes(raw_b64, nbytes, q):
    """Server side of the loop in a separate process: open the IPC handle, read the
    input, write input+1 back (CUDA IPC needs distinct processes, README.md:196-204)."""
    try:
        import base64 as b64
        import ctypes

        import numpy as np

        from client_b200 import _native

        lib = _native.load()
        ctx = _native.Context(0)
        raw = (ctypes.c_uint8 * 64).from_buffer_copy(b64.b64decode(raw_b64))
        region = ctypes.c_void_p()
        _native.check(lib.tb200_region_open(raw, nbytes, 0, ctypes.byref(region)))
        host = np.zeros(nbytes, np.uint8)
        _native.check(lib.tb200_region_read_host(ctx.handle, region, 0, host.ctypes.data, nbytes))
        vals = host.view(np.int32) + 1
        _native.check(lib.tb200_region_write_host(ctx.handle, region, 0, vals.ctypes.data, nbytes))
        lib.tb200_region_destroy(region)
        q.put(("ok", int(host.view(np.int32).sum())))
    except Exception as ex:  # pragma: no cover
        q.put(("error", repr(ex)))


def test_ipc_handle_opens_in_another_process(cudashm):
    a = np.arange(1024, dtype=np.int32)
    h = cudashm.create_shared_memory_region("ipc", a.nbytes, 0)
    cudashm.set_shared_memory_region(h, [a])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_peer_reads_and_writes, args=(cudashm.get_raw_handle(h), a.nbytes, q))
    p.start()
    status, val = q.get(timeout=120)
    p.join(60)
    assert status == "ok", val
    assert val == int(a.sum())
    assert np.array_equal(cudashm.get_contents_as_numpy(h, np.int32, [1024]), a + 1)
    cudashm.destroy_shared_memory_region(h)


def test_device_side_producers(cudashm):
    from oracle import cref

    h = cudashm.create_shared_memory_region("slot0", 602112 + 4000, 0)
    cudashm.fill_shared_memory_region(h, "FP32", [3, 224, 224], seed=7, stream_id=11)
    got = cudashm.get_contents_as_numpy(h, np.float32, [3, 224, 224])
    assert np.array_equal(got.view(np.uint8).reshape(-1), cref.fill(602112, "FP32", seed=7, stream=11))
    assert got.min() >= 0.0 and got.max() < 1.0
    res = cudashm.check_shared_memory_region(h, "sum", byte_size=602112)
    assert (res["sum"], res["xor32"]) == cref.checksum(got)
    img = np.random.default_rng(1).integers(0, 256, (224, 224, 3), dtype=np.uint8)
    cudashm.set_shared_memory_region_from_image(h, img, "FP32", "INCEPTION")
    got = cudashm.get_contents_as_numpy(h, np.float32, [3, 224, 224])
    ref = ((img.astype(np.float32) / 127.5) - 1).transpose(2, 0, 1)
    assert np.array_equal(got, ref)
    cudashm.destroy_shared_memory_region(h)


@pytest.mark.parametrize("datatype", ["FP32", "FP16", "BF16"])
def test_topk_on_device_matches_oracle(cudashm, datatype):
    """tb200_topk_async vs the oracle (numpy stable argsort order): ties, signed zeros,
    NaN, +-inf, k larger than the vector, several vectors per launch."""
    from client_b200 import _native
    from client_b200.device import DeviceOps
    from client_b200.utils import serialize_bf16_tensor
    from oracle import cref

    rng = np.random.default_rng(5)
    ops = DeviceOps(_native.default_context(0))
    lengths = [1, 7, 1000, 1000, 4096, 33]
    vecs = []
    for n in lengths:
        x = rng.standard_normal(n).astype(np.float32)
        x[rng.random(n) < 0.2] = np.float32(0.5)      # ties
        x[rng.random(n) < 0.05] = np.nan
        x[rng.random(n) < 0.05] = -0.0
        x[rng.random(n) < 0.02] = np.inf
        x[rng.random(n) < 0.02] = -np.inf
        if datatype == "FP16":
            x = x.astype(np.float16).astype(np.float32)
        elif datatype == "BF16":
            x = (x.view(np.uint32) & 0xFFFF0000).view(np.float32)
        vecs.append(x)
    es = 4 if datatype == "FP32" else 2
    h = cudashm.create_shared_memory_region("topk_data", sum(lengths) * 4 + 16 * len(lengths), 0)
    jobs, off, keep = [], 0, []
    for x in vecs:
        if datatype == "FP32":
            raw = x
        elif datatype == "FP16":
            raw = x.astype(np.float16)
        else:
            raw = np.frombuffer(serialize_bf16_tensor(x).item(), dtype=np.uint16)
        raw = np.ascontiguousarray(raw)
        keep.append(raw)
        ops.h2d(h._base_addr + off, raw.ctypes.data, raw.nbytes)
        jobs.append((h._base_addr + off, x.size, datatype))
        off += (x.size * es + 15) // 16 * 16
    ops.sync()
    for k in (1, 5, 40):
        values, indices = ops.topk(jobs, k)
        for j, x in enumerate(vecs):
            want_v, want_i = cref.topk(x, k)
            assert np.array_equal(indices[j], want_i), (datatype, k, j)
            assert np.array_equal(values[j], want_v, equal_nan=True)
    cudashm.destroy_shared_memory_region(h)


def test_classify_shared_memory_region(cudashm):
    """Client-side stand-in for the server's classification extension on outputs that stay
    in shared memory: b"<value>:<index>[:<label>]" like InferResult.as_numpy returns."""
    logits = np.zeros((2, 1000), dtype=np.float32)
    logits[0, [3, 500, 999]] = [2.5, 9.25, 2.5]
    logits[1, [7, 8]] = [-1.0, 4.0]
    h = cudashm.create_shared_memory_region("cls_data", logits.nbytes, 0)
    cudashm.set_shared_memory_region(h, [logits])
    got = cudashm.classify_shared_memory_region(h, "FP32", [2, 1000], 3)
    assert got.shape == (2, 3) and got.dtype == np.object_
    assert [g.decode() for g in got[0]] == ["9.250000:500", "2.500000:3", "2.500000:999"]
    assert [g.decode() for g in got[1]] == ["4.000000:8", "0.000000:0", "0.000000:1"]
    labels = ["c%d" % i for i in range(1000)]
    got = cudashm.classify_shared_memory_region(h, "FP32", [2, 1000], 1, labels=labels)
    assert got[0, 0] == b"9.250000:500:c500"
    cls = got[1, 0].decode().split(":")           # image_client.postprocess parsing
    assert (float(cls[0]), int(cls[1]), cls[2]) == (4.0, 8, "c8")
    with pytest.raises(cudashm.CudaSharedMemoryException):
        cudashm.classify_shared_memory_region(h, "INT32", [2, 1000], 1)
    with pytest.raises(cudashm.CudaSharedMemoryException):
        cudashm.classify_shared_memory_region(h, "FP32", [3, 1000], 1)
    cudashm.destroy_shared_memory_region(h)
