"""CUDA shared memory end to end: the client fills / packs tensors inside CUDA-IPC
regions on the device, registers the regions with the mock server (its own process,
which opens the IPC handles), sends requests that only name the regions, and
validates the outputs on the device.  Restates
src/python/examples/simple_http_cudashm_client.py:82-195 and simple_grpc_cudashm_client.py."""

import numpy as np
import pytest

from test_loopback import start_server

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def server():
    proc, http_port, grpc_port = start_server()
    yield {"http": "127.0.0.1:%d" % http_port, "grpc": "127.0.0.1:%d" % grpc_port}
    proc.terminate()
    proc.wait(10)


def _simple_regions(cudashm):
    ip = cudashm.create_shared_memory_region("input_data", 128, 0)
    op = cudashm.create_shared_memory_region("output_data", 128, 0)
    return ip, op


def test_http_cudashm_simple(server):
    import client_b200.http as httpclient
    import client_b200.utils.cuda_shared_memory as cudashm

    a = np.arange(16, dtype=np.int32)
    b = np.ones(16, dtype=np.int32)
    ip, op = _simple_regions(cudashm)
    with httpclient.InferenceServerClient(server["http"]) as client:
        client.unregister_cuda_shared_memory()
        cudashm.set_shared_memory_region(ip, [a, b])
        client.register_cuda_shared_memory("input_data", cudashm.get_raw_handle(ip), 0, 128)
        client.register_cuda_shared_memory("output_data", cudashm.get_raw_handle(op), 0, 128)
        assert sorted(r["name"] for r in client.get_cuda_shared_memory_status()) == ["input_data", "output_data"]
        inputs = [httpclient.InferInput("INPUT0", [1, 16], "INT32"), httpclient.InferInput("INPUT1", [1, 16], "INT32")]
        inputs[0].set_shared_memory("input_data", 64)
        inputs[1].set_shared_memory("input_data", 64, offset=64)
        outputs = [httpclient.InferRequestedOutput("OUTPUT0"), httpclient.InferRequestedOutput("OUTPUT1")]
        outputs[0].set_shared_memory("output_data", 64)
        outputs[1].set_shared_memory("output_data", 64, offset=64)
        results = client.infer("simple", inputs, outputs=outputs)
        assert results.get_output("OUTPUT0")["parameters"]["shared_memory_byte_size"] == 64
        got = cudashm.get_contents_as_numpy(op, np.int32, [2, 16])
        assert np.array_equal(got[0], a + b) and np.array_equal(got[1], a - b)
        # and on the device, without bringing the outputs to the host
        from client_b200.device import DeviceOps

        res = DeviceOps(device_id=0).check_one("addsub", op._base_addr, 64, b=op._base_addr + 64, c=ip._base_addr, d=ip._base_addr + 64)
        assert res["mismatches"] == 0
        client.unregister_cuda_shared_memory("input_data")
        client.unregister_cuda_shared_memory()
    cudashm.destroy_shared_memory_region(ip)
    cudashm.destroy_shared_memory_region(op)


def test_grpc_cudashm_device_generated_inputs(server):
    """C2 shape over gRPC: the input is generated on the device inside the region,
    the server computes from the IPC mapping, the output lands in an output region."""
    import client_b200.grpc as grpcclient
    import client_b200.utils.cuda_shared_memory as cudashm
    from oracle import cref

    n_in, n_out = 602112, 4000
    ip = cudashm.create_shared_memory_region("dn_in", n_in, 0)
    op = cudashm.create_shared_memory_region("dn_out", n_out, 0)
    with grpcclient.InferenceServerClient(server["grpc"]) as client:
        client.unregister_cuda_shared_memory()
        client.register_cuda_shared_memory("dn_in", cudashm.get_raw_handle(ip), 0, n_in)
        client.register_cuda_shared_memory("dn_out", cudashm.get_raw_handle(op), 0, n_out)
        assert set(client.get_cuda_shared_memory_status().regions) == {"dn_in", "dn_out"}
        inp = grpcclient.InferInput("data_0", [3, 224, 224], "FP32").set_shared_memory("dn_in", n_in)
        out = grpcclient.InferRequestedOutput("fc6_1")
        out.set_shared_memory("dn_out", n_out)
        for request in range(3):
            cudashm.fill_shared_memory_region(ip, "FP32", [3, 224, 224], seed=5, stream_id=request)
            client.infer("densenet_onnx", [inp], outputs=[out])
            x = cref.fill(n_in, "FP32", seed=5, stream=request).view(np.float32)
            pad = (-x.size) % 1000
            want = np.concatenate([x, np.zeros(pad, np.float32)]).reshape(-1, 1000).sum(axis=0, dtype=np.float32) / np.float32(151)
            got = cudashm.get_contents_as_numpy(op, np.float32, [1000])
            assert np.array_equal(got, want)
            res = cudashm.check_shared_memory_region(op, "top1", byte_size=n_out)
            assert res["argmax"] == int(np.argmax(want)) and res["mismatches"] == 0
        client.unregister_cuda_shared_memory()
    cudashm.destroy_shared_memory_region(ip)
    cudashm.destroy_shared_memory_region(op)


@pytest.mark.parametrize("algorithm", ["gzip", "deflate"])
def test_request_body_compressed_on_the_device(server, algorithm):
    """request_compression_algorithm with the device encoder switched on: the mock server
    inflates the body with zlib / gzip (like the reference's peer) and computes add/sub."""
    import client_b200.http as httpclient

    a = np.arange(16, dtype=np.int32)[None, :]
    b = np.full((1, 16), 7, dtype=np.int32)
    big = np.zeros((1, 1 << 18), dtype=np.int32)  # 1 MB of mostly-zero data through identity
    big[0, ::977] = np.arange(big[0, ::977].size)
    httpclient.set_device_compression(True)
    try:
        with httpclient.InferenceServerClient(server["http"]) as client:
            inputs = [httpclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a),
                      httpclient.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(b)]
            res = client.infer("simple", inputs, request_compression_algorithm=algorithm)
            assert np.array_equal(res.as_numpy("OUTPUT0"), a + b) and np.array_equal(res.as_numpy("OUTPUT1"), a - b)
            inp = httpclient.InferInput("INPUT0", list(big.shape), "INT32").set_data_from_numpy(big)
            res = client.infer("custom_identity_int32", [inp], request_compression_algorithm=algorithm,
                               response_compression_algorithm=algorithm)
            assert np.array_equal(res.as_numpy("OUTPUT0"), big)
    finally:
        httpclient.set_device_compression(False)
