"""asyncio clients against the mock server (reference examples
simple_http_aio_infer_client.py / simple_grpc_aio_infer_client.py /
simple_grpc_aio_sequence_stream_infer_client.py restated)."""

import asyncio

import numpy as np
import pytest

from client_b200.utils import InferenceServerException
from test_loopback import start_server


@pytest.fixture(scope="module")
def server():
    proc, http_port, grpc_port = start_server()
    yield {"http": "127.0.0.1:%d" % http_port, "grpc": "127.0.0.1:%d" % grpc_port}
    proc.terminate()
    proc.wait(10)


def test_http_aio(server):
    import client_b200.http.aio as aioclient

    async def run():
        async with aioclient.InferenceServerClient(server["http"]) as client:
            assert await client.is_server_live() and await client.is_model_ready("simple")
            assert (await client.get_model_metadata("simple"))["name"] == "simple"
            a = np.arange(16, dtype=np.int32)[None, :]
            b = np.ones((1, 16), dtype=np.int32)
            inputs = [aioclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a),
                      aioclient.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(b, binary_data=False)]
            outputs = [aioclient.InferRequestedOutput("OUTPUT0"), aioclient.InferRequestedOutput("OUTPUT1", binary_data=False)]
            results = await asyncio.gather(*[client.infer("simple", inputs, outputs=outputs, request_id=str(i)) for i in range(8)])
            for r in results:
                assert np.array_equal(r.as_numpy("OUTPUT0"), a + b) and np.array_equal(r.as_numpy("OUTPUT1"), a - b)
            r = await client.infer("simple", inputs, request_compression_algorithm="deflate", response_compression_algorithm="gzip")
            assert np.array_equal(r.as_numpy("OUTPUT0"), a + b)
            with pytest.raises(InferenceServerException):
                await client.infer("nope", inputs)

    asyncio.run(run())


def test_grpc_aio_and_stream(server):
    import client_b200.grpc.aio as aioclient

    async def run():
        async with aioclient.InferenceServerClient(server["grpc"]) as client:
            assert await client.is_server_ready() and await client.is_model_ready("simple")
            assert (await client.get_model_metadata("simple", as_json=True))["name"] == "simple"
            a = np.arange(16, dtype=np.int32)[None, :]
            b = np.ones((1, 16), dtype=np.int32)
            inputs = [aioclient.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a),
                      aioclient.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(b)]
            results = await asyncio.gather(*[client.infer("simple", inputs) for _ in range(8)])
            for r in results:
                assert np.array_equal(r.as_numpy("OUTPUT0"), a + b)
            with pytest.raises(InferenceServerException):
                await client.infer("nope", inputs)

            async def requests():
                tok = np.arange(64, dtype=np.int32).reshape(1, 64)
                yield {"model_name": "llama3_8b", "inputs": [aioclient.InferInput("input_ids", [1, 64], "INT32").set_data_from_numpy(tok)],
                       "request_id": "a", "parameters": {"max_tokens": 3}}
                yield {"model_name": "nope", "inputs": []}

            got, errs = [], []
            async for result, error in client.stream_infer(requests()):
                (errs if error is not None else got).append(result if error is None else error)
            assert len(got) == 3 and len(errs) == 1
            base = int(np.arange(64).sum() % 128256)
            assert [int(r.as_numpy("token")[0, 0]) for r in got] == [base, base + 1, base + 2]

    asyncio.run(run())
