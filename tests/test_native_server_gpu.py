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
