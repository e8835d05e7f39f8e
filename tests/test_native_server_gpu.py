"""The native stand-in server (csrc/mock_server.cu, own process) driven by the unchanged
client API over CUDA shared memory, and by the native load generator.  Flow of
src/python/examples/simple_http_cudashm_client.py:82-195."""

import os
import subprocess
import sys

import numpy as np
import pytest

from client_b200.perf import cli

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def server():
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    proc = subprocess.Popen([sys.executable, "-m", "client_b200.testing.native_server", "--port", str(port)],
                            cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    line = proc.stdout.readline()
    assert "listening" in line, line + proc.stdout.read()
    yield "127.0.0.1:%d" % port
    proc.terminate()
    proc.wait(10)


def test_simple_and_densenet_over_cuda_shm(server):
    import client_b200.http as httpclient
    import client_b200.utils.cuda_shared_memory as cudashm

    a = np.arange(16, dtype=np.int32)
    b = np.full(16, 3, dtype=np.int32)
    ip = cudashm.create_shared_memory_region("ns_in", 128, 0)
    op = cudashm.create_shared_memory_region("ns_out", 128, 0)
    din = cudashm.create_shared_memory_region("ns_data", 602112, 0)
    dout = cudashm.create_shared_memory_region("ns_fc6", 4000, 0)
    with httpclient.InferenceServerClient(server) as client:
        assert client.is_server_live() and client.is_model_ready("simple")
        assert client.get_model_metadata("densenet_onnx")["inputs"][0]["shape"] == [3, 224, 224]
        cudashm.set_shared_memory_region(ip, [a, b])
        for name, h, n in (("ns_in", ip, 128), ("ns_out", op, 128), ("ns_data", din, 602112), ("ns_fc6", dout, 4000)):
            client.register_cuda_shared_memory(name, cudashm.get_raw_handle(h), 0, n)
        assert sorted(r["name"] for r in client.get_cuda_shared_memory_status()) == ["ns_data", "ns_fc6", "ns_in", "ns_out"]
        with pytest.raises(Exception, match="already in manager"):
            client.register_cuda_shared_memory("ns_in", cudashm.get_raw_handle(ip), 0, 128)
        inputs = [httpclient.InferInput("INPUT0", [1, 16], "INT32"), httpclient.InferInput("INPUT1", [1, 16], "INT32")]
        inputs[0].set_shared_memory("ns_in", 64)
        inputs[1].set_shared_memory("ns_in", 64, offset=64)
        outputs = [httpclient.InferRequestedOutput("OUTPUT0"), httpclient.InferRequestedOutput("OUTPUT1")]
        outputs[0].set_shared_memory("ns_out", 64)
        outputs[1].set_shared_memory("ns_out", 64, offset=64)
        res = client.infer("simple", inputs, outputs=outputs)
        assert res.get_output("OUTPUT1")["parameters"]["shared_memory_byte_size"] == 64
        got = cudashm.get_contents_as_numpy(op, np.int32, [2, 16])
        assert np.array_equal(got[0], a + b) and np.array_equal(got[1], a - b)

        cudashm.fill_shared_memory_region(din, "FP32", [3, 224, 224], seed=5, stream_id=2)
        x = cudashm.get_contents_as_numpy(din, np.float32, [150528]).copy()
        di = httpclient.InferInput("data_0", [3, 224, 224], "FP32").set_shared_memory("ns_data", 602112)
        do = httpclient.InferRequestedOutput("fc6_1")
        do.set_shared_memory("ns_fc6", 4000)
        client.infer("densenet_onnx", [di], outputs=[do])
        y = cudashm.get_contents_as_numpy(dout, np.float32, [1000])
        xs = np.zeros(151 * 1000, dtype=np.float32)
        xs[:150528] = x
        want = xs.reshape(151, 1000).sum(axis=0, dtype=np.float32) / np.float32(151)
        assert np.array_equal(y, want)
        with pytest.raises(Exception, match="unknown model"):
            client.infer("nope", [di], outputs=[do])
        client.unregister_cuda_shared_memory("ns_in")
        client.unregister_cuda_shared_memory()
        assert client.get_cuda_shared_memory_status() == []
    for h in (ip, op, din, dout):
        cudashm.destroy_shared_memory_region(h)


def test_native_engine_against_native_server(server):
    rows = cli.main(["-m", "densenet_onnx", "-u", server, "--shared-memory", "cuda", "--engine", "native",
                     "--concurrency-range", "4:16:4x", "-p", "300", "-r", "3", "--json"])
    assert [r["concurrency"] for r in rows] == [4, 16]
    for r in rows:
        assert r["count"] > 50 and r["failed"] == 0 and r["nonfinite"] == 0, r
    rows = cli.main(["-m", "densenet_onnx", "-u", server, "--shared-memory", "cuda", "--concurrency-range", "2",
                     "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 10 and rows[0]["failed"] == 0 and rows[0]["nonfinite"] == 0


def test_reference_cc_cudashm_example_if_prebuilt(server):
    """The reference's simple_http_cudashm_client.cc (cudaMalloc + cudaIpcGetMemHandle +
    RegisterCudaSharedMemory by hand), compiled unmodified against the C++ front end in the
    build container (oracle/build_ref_examples.py -> oracle/_ref/), run here against the
    native server."""
    exe = os.path.join(ROOT, "oracle", "_ref", "cc_examples", "simple_http_cudashm_client")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/cc_examples was not prebuilt")
    r = subprocess.run([exe, "-u", server], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PASS : Cuda Shared Memory" in r.stdout, r.stdout[-800:] + r.stderr[-400:]


# ---- the gRPC side of the native server (csrc/grpc_server.h + generated message classes) --------
@pytest.fixture(scope="module")
def grpc_server():
    proc = subprocess.Popen([sys.executable, "-m", "client_b200.testing.native_server", "--port", "0", "--grpc-port", "0"],
                            cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    line = proc.stdout.readline()
    assert "listening" in line and "grpc=" in line, line + proc.stdout.read()
    yield {"http": line.split()[5], "grpc": line.split("grpc=")[1].strip()}
    proc.terminate()
    proc.wait(10)


def test_grpc_models_with_tensors_in_the_message(grpc_server):
    """grpcio client -> native server: health / metadata / config RPCs, `simple` and `bert_large`
    with raw_input_contents (the model runs as a CUDA kernel on pinned slabs), values against numpy
    (the Python mock's formulas), errors as grpc statuses."""
    import client_b200.grpc as grpcclient

    with grpcclient.InferenceServerClient(grpc_server["grpc"]) as client:
        assert client.is_server_live() and client.is_server_ready() and client.is_model_ready("bert_large") and not client.is_model_ready("nope")
        assert client.get_server_metadata().name == "triton"
        md = client.get_model_metadata("bert_large")
        assert [(i.name, i.datatype, list(i.shape)) for i in md.inputs] == [("input_ids", "INT64", [1, 384]), ("attention_mask", "INT64", [1, 384])]
        cfg = client.get_model_config("llama3_8b").config
        assert cfg.model_transaction_policy.decoupled and cfg.input[0].dims == [1, -1]
        assert {m.name for m in client.get_model_repository_index().models} >= {"densenet_onnx", "simple", "bert_large", "llama3_8b"}
        with pytest.raises(Exception, match="unknown model"):
            client.get_model_metadata("nope")
        a = np.arange(16, dtype=np.int32)[None, :]
        b = np.full((1, 16), 5, np.int32)
        ins = [grpcclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a), grpcclient.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(b)]
        res = client.infer("simple", ins, request_id="r1")
        assert np.array_equal(res.as_numpy("OUTPUT0"), a + b) and np.array_equal(res.as_numpy("OUTPUT1"), a - b) and res.get_response().id == "r1"
        rng = np.random.default_rng(3)
        for _ in range(5):
            ids = rng.integers(0, 30522, (1, 384), dtype=np.int64)
            mask = rng.integers(0, 2, (1, 384), dtype=np.int64)
            ins = [grpcclient.InferInput("input_ids", [1, 384], "INT64").set_data_from_numpy(ids),
                   grpcclient.InferInput("attention_mask", [1, 384], "INT64").set_data_from_numpy(mask)]
            got = client.infer("bert_large", ins).as_numpy("logits")
            want = ((ids % 1000).astype(np.float32) * mask.astype(np.float32)) / np.float32(1000)
            assert got.dtype == np.float32 and got.shape == (1, 384) and np.array_equal(got, want)  # same IEEE operations: bit-exact
        with pytest.raises(Exception, match="unknown model"):
            client.infer("nope", ins)
        with pytest.raises(Exception, match="bytes, shape needs"):
            bad = [grpcclient.InferInput("input_ids", [1, 384], "INT64").set_data_from_numpy(ids), grpcclient.InferInput("attention_mask", [1, 100], "INT64").set_data_from_numpy(mask[:, :100])]
            bad[1]._input.ClearField("shape")
            bad[1]._input.shape.extend([1, 384])
            client.infer("bert_large", bad)
        with pytest.raises(Exception, match="decoupled"):
            client.infer("llama3_8b", [grpcclient.InferInput("input_ids", [1, 8], "INT32").set_data_from_numpy(np.ones((1, 8), np.int32))])


def test_grpc_stream_and_cuda_shm(grpc_server):
    """ModelStreamInfer on the native server: the decoupled llama3_8b answers max_tokens times
    (token k = (sum(ids) + k) mod 128256, the Python mock's rule), a plain model once, an unknown
    model through error_message; densenet_onnx over gRPC with CUDA shared memory."""
    import threading

    import client_b200.grpc as grpcclient
    import client_b200.utils.cuda_shared_memory as cudashm

    with grpcclient.InferenceServerClient(grpc_server["grpc"]) as client:
        got, cond = [], threading.Condition()

        def on_response(result, error):
            with cond:
                got.append((result, error))
                cond.notify_all()

        client.start_stream(callback=on_response)
        ids = np.random.default_rng(5).integers(0, 128256, (1, 4096), dtype=np.int32)
        inp = grpcclient.InferInput("input_ids", [1, 4096], "INT32").set_data_from_numpy(ids)
        client.async_stream_infer("llama3_8b", [inp], request_id="p0", parameters={"max_tokens": 6})
        with cond:
            assert cond.wait_for(lambda: len(got) == 6, 20)
        assert all(e is None for _, e in got)
        base = int(ids.astype(np.int64).sum() % 128256)
        assert [int(r.as_numpy("token")[0, 0]) for r, _ in got] == [(base + k) % 128256 for k in range(6)]
        assert [r.get_response().parameters["triton_final_response"].bool_param for r, _ in got] == [False] * 5 + [True]
        got.clear()
        client.async_stream_infer("llama3_8b", [inp], parameters={"max_tokens": 2}, enable_empty_final_response=True)
        with cond:
            assert cond.wait_for(lambda: len(got) == 3, 20)
        assert [r.get_response().parameters["triton_final_response"].bool_param for r, _ in got] == [False, False, True]
        assert len(got[2][0].get_response().outputs) == 0
        got.clear()
        a = np.arange(16, dtype=np.int32)[None, :]
        ins = [grpcclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a), grpcclient.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(a)]
        client.async_stream_infer("simple", ins)
        client.async_stream_infer("nope", ins)
        with cond:
            assert cond.wait_for(lambda: len(got) == 2, 20)
        assert got[0][1] is None and np.array_equal(got[0][0].as_numpy("OUTPUT0"), 2 * a)
        assert got[1][1] is not None and "unknown model" in str(got[1][1])
        client.stop_stream()
        # cuda shared memory over gRPC
        din = cudashm.create_shared_memory_region("ng_data", 602112, 0)
        dout = cudashm.create_shared_memory_region("ng_fc6", 4000, 0)
        x = np.random.default_rng(1).random((3, 224, 224), dtype=np.float32)
        cudashm.set_shared_memory_region(din, [x])
        client.register_cuda_shared_memory("ng_data", cudashm.get_raw_handle(din), 0, 602112)
        client.register_cuda_shared_memory("ng_fc6", cudashm.get_raw_handle(dout), 0, 4000)
        assert sorted(client.get_cuda_shared_memory_status().regions) == ["ng_data", "ng_fc6"]
        i0 = grpcclient.InferInput("data_0", [3, 224, 224], "FP32")
        i0.set_shared_memory("ng_data", 602112)
        o0 = grpcclient.InferRequestedOutput("fc6_1")
        o0.set_shared_memory("ng_fc6", 4000)
        client.infer("densenet_onnx", [i0], outputs=[o0])
        got_out = cudashm.get_contents_as_numpy(dout, np.float32, [1000])
        flat = x.reshape(-1)
        xs = np.concatenate([flat, np.zeros((-flat.size) % 1000, np.float32)]).reshape(-1, 1000)
        assert np.allclose(got_out, xs.sum(axis=0, dtype=np.float32) / np.float32(xs.shape[0]), rtol=1e-5, atol=1e-6)
        client.unregister_cuda_shared_memory()
        assert len(client.get_cuda_shared_memory_status().regions) == 0
        cudashm.destroy_shared_memory_region(din)
        cudashm.destroy_shared_memory_region(dout)


def test_native_generator_against_native_grpc_server(grpc_server):
    """BASELINE configs[3] and [4] end to end without Python in the loop: --engine native over gRPC
    (tensors generated by the fill kernel into the message tails) against the native server running
    the models as CUDA kernels; C5 on ModelStreamInfer streams with TTFT."""
    rows = cli.main(["-m", "bert_large", "-u", grpc_server["grpc"], "-i", "grpc", "--shared-memory", "none", "--engine", "native",
                     "--concurrency-range", "16", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 100 and rows[0]["failed"] == 0 and rows[0]["input_bytes"] == 6144, rows
    rows = cli.main(["-m", "llama3_8b", "-u", grpc_server["grpc"], "-i", "grpc", "--streaming", "--engine", "native", "--shape", "input_ids:1,4096",
                     "--shared-memory", "none", "--concurrency-range", "8", "-p", "300", "-r", "3", "--request-parameter", "max_tokens:8:int", "--json"])
    r = rows[0]
    assert r["count"] > 50 and r["failed"] == 0 and r["responses"] == 8 * r["count"] and 0 < r["ttft_p50_us"] <= r["p50_us"], r
    rows = cli.main(["-m", "densenet_onnx", "-u", grpc_server["grpc"], "-i", "grpc", "--shared-memory", "cuda", "--engine", "native",
                     "--concurrency-range", "8", "-p", "300", "-r", "3", "--json"])
    assert rows[0]["count"] > 50 and rows[0]["failed"] == 0 and rows[0]["nonfinite"] == 0, rows


def test_reference_cc_grpc_examples_against_native_server_if_prebuilt(grpc_server):
    """simple_grpc_infer_client.cc / simple_grpc_async_infer_client.cc (reference sources, compiled
    unmodified against the C++ gRPC front end) against the native server: C++ on both ends."""
    for name, mark in (("simple_grpc_infer_client", "PASS : Infer"), ("simple_grpc_async_infer_client", "PASS : Async Infer")):
        exe = os.path.join(ROOT, "oracle", "_ref", "cc_examples", name)
        if not os.path.exists(exe):
            pytest.skip("oracle/_ref/cc_examples was not prebuilt")
        r = subprocess.run([exe, "-u", grpc_server["grpc"]], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0 and mark in r.stdout, r.stdout[-800:] + r.stderr[-400:]


def test_lookahead_over_cuda_shm(server):
    """--lookahead 4 with CUDA shared memory: four input / output images per slot, requests name
    the image's offsets, one device pass validates four outputs and regenerates four inputs."""
    rows = cli.main(["-m", "densenet_onnx", "-u", server, "--shared-memory", "cuda", "--engine", "native", "--lookahead", "4",
                     "--concurrency-range", "1:8:8x", "-p", "300", "-r", "3", "--json"])
    assert [r["concurrency"] for r in rows] == [1, 8]
    for r in rows:
        assert r["count"] > 20 and r["failed"] == 0 and r["nonfinite"] == 0, r
        assert 0 < r["device_slots"] <= r["count"] / 4 + 8, r


@pytest.mark.parametrize("mode", ["device", "per-request", "once"])
def test_host_loop_modes(server, mode):
    """bench.py's e2e loops (client_b200/perf/host_loop.py): every mode completes validated requests"""
    from client_b200.perf import host_loop

    n, lat = host_loop.run_loop(server, 0, "t_" + mode.replace("-", ""), 0.3, mode)
    assert n > 10 and len(lat) == n


def test_reference_loop_same_server(server):
    """the reference arm of bench.py (oracle/ref_client.py, cuda-python) against the same server"""
    pytest.importorskip("cuda.bindings.runtime")
    from oracle import ref_client

    n, lat = ref_client.run_loop(server, 0, "t_ref", 0.3, "per-request")
    assert n > 5 and len(lat) == n
