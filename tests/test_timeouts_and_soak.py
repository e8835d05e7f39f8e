"""Client-side deadlines and a memory-growth soak, after the reference's
src/c++/tests/client_timeout_test.cc:212-340 (sync / async / C++ "Deadline Exceeded") and
src/python/examples/memory_growth_test.py:73-113 (new client per iteration, output == input)."""

import os
import resource
import subprocess

import numpy as np
import pytest

import client_b200.grpc as grpcclient
import client_b200.http as httpclient
from client_b200.utils import InferenceServerException
from test_loopback import start_server

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def slow_server():
    proc, http_port, grpc_port = start_server(extra=("--delay-us", "300000"))  # 0.3 s per inference
    yield {"http": "127.0.0.1:%d" % http_port, "grpc": "127.0.0.1:%d" % grpc_port}
    proc.terminate()
    proc.wait(10)


def _simple_inputs(mod):
    a = np.arange(16, dtype=np.int32)[None, :]
    b = np.ones((1, 16), dtype=np.int32)
    return [mod.InferInput("INPUT0", [1, 16], "INT32").set_data_from_numpy(a),
            mod.InferInput("INPUT1", [1, 16], "INT32").set_data_from_numpy(b)], a, b


def test_grpc_client_timeout_sync_and_async(slow_server):
    with grpcclient.InferenceServerClient(slow_server["grpc"]) as client:
        inputs, a, b = _simple_inputs(grpcclient)
        with pytest.raises(InferenceServerException) as info:
            client.infer("simple", inputs, client_timeout=0.05)
        assert "DEADLINE_EXCEEDED" in info.value.status() and "Deadline Exceeded" in info.value.message()
        # a generous deadline succeeds
        res = client.infer("simple", inputs, client_timeout=5.0)
        assert np.array_equal(res.as_numpy("OUTPUT0"), a + b)
        # async: the error reaches the callback
        got = []
        import threading

        done = threading.Event()

        def cb(result, error):
            got.append((result, error))
            done.set()

        client.async_infer("simple", inputs, cb, client_timeout=0.05)
        assert done.wait(10)
        assert got[0][0] is None and "DEADLINE_EXCEEDED" in got[0][1].status()


def test_http_network_timeout(slow_server):
    # the Python HTTP client's deadline is the connection's network_timeout (reference
    # http/_client.py:158-170): the socket times out, nothing is swallowed
    with httpclient.InferenceServerClient(slow_server["http"], network_timeout=0.05) as client:
        inputs, _, _ = _simple_inputs(httpclient)
        with pytest.raises((TimeoutError, OSError)):
            client.infer("simple", inputs)
    with httpclient.InferenceServerClient(slow_server["http"], network_timeout=5.0) as client:
        inputs, a, b = _simple_inputs(httpclient)
        assert np.array_equal(client.infer("simple", inputs).as_numpy("OUTPUT1"), a - b)


def test_cc_client_deadline_exceeded(slow_server):
    """C++ front end: InferOptions::client_timeout_ (microseconds) -> "Deadline Exceeded"
    (http_client.cc:1814-1817)."""
    from client_b200.build import build_cpp_client, build_native

    build_native()
    build_cpp_client()
    src = os.path.join(ROOT, "build", "cc_timeout.cc")
    exe = os.path.join(ROOT, "build", "cc_timeout")
    with open(src, "w") as fh:
        fh.write(r'''
#include <iostream>
#include "http_client.h"
namespace tc = triton::client;
int main(int argc, char** argv) {
  std::unique_ptr<tc::InferenceServerHttpClient> client;
  tc::InferenceServerHttpClient::Create(&client, argv[1]);
  int32_t a[16], b[16];
  for (int i = 0; i < 16; ++i) { a[i] = i; b[i] = 1; }
  tc::InferInput *in0, *in1;
  tc::InferInput::Create(&in0, "INPUT0", {1, 16}, "INT32");
  tc::InferInput::Create(&in1, "INPUT1", {1, 16}, "INT32");
  in0->AppendRaw(reinterpret_cast<uint8_t*>(a), 64);
  in1->AppendRaw(reinterpret_cast<uint8_t*>(b), 64);
  tc::InferOptions options("simple");
  options.client_timeout_ = 50000;  // 50 ms against a 300 ms model
  tc::InferResult* result = nullptr;
  tc::Error err = client->Infer(&result, options, {in0, in1});
  std::cout << "first: " << err.Message() << std::endl;
  if (err.IsOk() || err.Message().find("Deadline Exceeded") == std::string::npos) return 1;
  options.client_timeout_ = 5000000;
  err = client->Infer(&result, options, {in0, in1});
  std::cout << "second: " << (err.IsOk() ? "ok" : err.Message()) << std::endl;
  return err.IsOk() ? 0 : 2;
}
''')
    cpp = os.path.join(ROOT, "client_b200", "cpp")
    libdir = os.path.join(ROOT, "client_b200", "lib")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(cpp, "compat"), "-I" + cpp, src, "-o", exe,
                    "-L" + libdir, "-ltb200client", "-ltb200", "-Wl,-rpath," + libdir, "-lpthread"], check=True)
    r = subprocess.run([exe, slow_server["http"]], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr


def test_memory_growth_new_client_per_iteration():
    """memory_growth_test.py: a fresh client per request against custom_identity_int32; the
    process must not grow (here: < 20 MB over 300 iterations after a warm-up)."""
    proc, http_port, grpc_port = start_server()
    try:
        def one(i):
            x = np.full((1, 8), i, dtype=np.int32)
            with httpclient.InferenceServerClient("127.0.0.1:%d" % http_port) as c:
                inp = httpclient.InferInput("INPUT0", [1, 8], "INT32").set_data_from_numpy(x)
                assert np.array_equal(c.infer("custom_identity_int32", [inp]).as_numpy("OUTPUT0"), x)
            with grpcclient.InferenceServerClient("127.0.0.1:%d" % grpc_port) as c:
                inp = grpcclient.InferInput("INPUT0", [1, 8], "INT32").set_data_from_numpy(x)
                assert np.array_equal(c.infer("custom_identity_int32", [inp]).as_numpy("OUTPUT0"), x)

        for i in range(50):
            one(i)
        before = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss
        for i in range(300):
            one(i)
        grown_kb = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss - before
        assert grown_kb < 20 * 1024, "process grew by %d KB" % grown_kb
    finally:
        proc.terminate()
        proc.wait(10)
