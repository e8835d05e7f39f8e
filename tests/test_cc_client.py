"""C++ front end (client_b200/cpp): compiles tests/cpp/test_cc_client.cc with
g++ and runs it -- the reference's HTTPJSONDataTest known answers
(src/c++/tests/cc_client_test.cc:1662-2170), scatter-list semantics, request-header layout,
and the InferMulti / AsyncInfer cases against the mock server's `simple` model over HTTP."""

import os
import subprocess

import pytest

from test_loopback import start_server

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "client_b200", "lib")


@pytest.fixture(scope="module")
def binary():
    from client_b200.build import build_cpp_client, build_native

    build_native()
    build_cpp_client()
    out = os.path.join(ROOT, "build", "test_cc_client")
    src = os.path.join(ROOT, "tests", "cpp", "test_cc_client.cc")
    cpp = os.path.join(ROOT, "client_b200", "cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-I" + os.path.join(cpp, "compat"), "-I" + cpp, src, "-o", out,
                    "-L" + LIBDIR, "-ltb200client", "-ltb200", "-Wl,-rpath," + LIBDIR, "-lpthread"], check=True)
    return out


def _compile(name):
    out = os.path.join(ROOT, "build", name)
    cpp = os.path.join(ROOT, "client_b200", "cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-Wall", "-I" + os.path.join(cpp, "compat"), "-I" + cpp,
                    os.path.join(ROOT, "tests", "cpp", name + ".cc"), "-o", out,
                    "-L" + LIBDIR, "-ltb200client", "-ltb200", "-Wl,-rpath," + LIBDIR, "-lpthread"], check=True)
    return out


def test_cudashm_example_compiles(binary):
    _compile("test_cc_cudashm")


@pytest.mark.gpu
def test_cudashm_flow_on_gpu(binary, tmp_path):
    """AppendRaw scatter list -> IPC region, request by region name against the native
    server (own process), device-side add/sub validation, device fill vs the oracle."""
    import socket
    import sys

    import numpy as np

    from oracle import cref

    exe = _compile("test_cc_cudashm")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    srv = subprocess.Popen([sys.executable, "-m", "client_b200.testing.native_server", "--port", str(port)],
                           cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    try:
        assert "listening" in srv.stdout.readline()
        dump = str(tmp_path / "fill.bin")
        r = subprocess.run([exe, "127.0.0.1:%d" % port, dump], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
        got = np.fromfile(dump, dtype=np.uint8)
        assert np.array_equal(got, cref.fill(602112, "FP32", seed=7, stream=3))
    finally:
        srv.terminate()
        srv.wait(10)


def test_known_answers_offline(binary):
    r = subprocess.run([binary], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PASS (offline)" in r.stdout, r.stdout + r.stderr


def test_loopback_simple_model(binary):
    proc, http_port, _ = start_server()
    try:
        r = subprocess.run([binary, "127.0.0.1:%d" % http_port], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS (offline + loopback)" in r.stdout, r.stdout + r.stderr
    finally:
        proc.terminate()
        proc.wait(10)


def _has_gpu():
    import ctypes

    from client_b200 import _native

    n = ctypes.c_int(0)
    return _native.load().tb200_device_count(ctypes.byref(n)) == 0 and n.value > 0


def test_body_compression_without_a_device(binary):
    """Responses: the mock server compresses on Accept-Encoding, zlib inflates.  Requests: made by the
    device encoder only -- without a CUDA device the call reports that."""
    if _has_gpu():
        pytest.skip("a GPU is present")
    proc, http_port, _ = start_server()
    try:
        r = subprocess.run([binary, "127.0.0.1:%d" % http_port, "compress-nogpu"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
    finally:
        proc.terminate()
        proc.wait(10)


@pytest.mark.gpu
def test_body_compression_on_the_device(binary):
    """gzip / deflate request bodies produced by tb200_deflate_async, inflated by the server; gzip responses."""
    proc, http_port, _ = start_server()
    try:
        r = subprocess.run([binary, "127.0.0.1:%d" % http_port, "compress-gpu"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and "PASS" in r.stdout, r.stdout + r.stderr
    finally:
        proc.terminate()
        proc.wait(10)
