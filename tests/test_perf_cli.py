"""The load generator's host logic end to end on the CPU box: system shared memory and
wire modes against the mock server.  The product generates request data on the device
and has no host fallback; this CPU-only test injects its own staging object (below) in
place of the pinned buffer + fill kernel.  The real device path is covered by
tests/test_perf_gpu.py."""

import numpy as np
import pytest

from client_b200.perf import cli
from client_b200.perf.loadgen import RequestRecord, summarize
from test_loopback import start_server


class HostStaging:
    """Test double for the pinned staging buffer + fill kernel."""

    def allocate(self, nbytes):
        self.buf = np.zeros(nbytes, np.uint8)

    def view(self, off, n):
        return memoryview(self.buf)[off:off + n]

    def fill(self, tensors, input_data, seed):
        rng = np.random.default_rng(seed)
        for off, t in tensors:
            if t.datatype == "BYTES":  # must be well-formed <u32 length><chars> elements
                from oracle import cref

                self.buf[off:off + t.nbytes] = cref.fill(t.nbytes, "BYTES", seed=seed, stream=off, irange=t.string_length)
                continue
            self.buf[off:off + t.nbytes] = 0 if input_data == "zero" else rng.integers(0, 256, t.nbytes, dtype=np.uint8)


@pytest.fixture(scope="module")
def server():
    proc, http_port, grpc_port = start_server()
    yield {"http": "127.0.0.1:%d" % http_port, "grpc": "127.0.0.1:%d" % grpc_port}
    proc.terminate()
    proc.wait(10)


def test_parse_range():
    assert cli.parse_range("4") == [4]
    assert cli.parse_range("1:4") == [1, 2, 3, 4]
    assert cli.parse_range("1:64:2x") == [1, 2, 4, 8, 16, 32, 64]
    assert cli.parse_range("2:8:3") == [2, 5, 8]


def test_summarize_percentiles():
    recs = [RequestRecord(0, (i + 1) * 1000, None, True) for i in range(100)] + [RequestRecord(0, 5, None, False)]
    s = summarize(recs, 2.0, percentile=95)
    assert s["count"] == 100 and s["failed"] == 1 and s["throughput"] == 50.0
    assert abs(s["p50_us"] - 50.5) < 1e-6 and abs(s["latency_us"] - np.percentile(np.arange(1, 101), 95)) < 1e-6


@pytest.mark.parametrize("protocol", ["http", "grpc"])
@pytest.mark.parametrize("shm", ["system", "none"])
def test_simple_model_closed_loop(server, protocol, shm):
    rows = cli.main(["-m", "simple", "-u", server[protocol], "-i", protocol, "--shared-memory", shm,
                     "--concurrency-range", "1:2", "-p", "200", "-r", "3", "--json"], staging_factory=HostStaging)
    assert [r["concurrency"] for r in rows] == [1, 2]
    for r in rows:
        assert r["count"] > 5 and r["failed"] == 0 and r["throughput"] > 10, r


def test_streaming_reports_ttft(server):
    rows = cli.main(["-m", "llama3_8b", "-u", server["grpc"], "-i", "grpc", "--streaming", "--shape", "input_ids:1,64",
                     "--concurrency-range", "2", "-p", "200", "-r", "3", "--request-parameter", "max_tokens:3:int", "--json"], staging_factory=HostStaging)
    assert rows[0]["count"] > 3 and "ttft_p50_us" in rows[0] and rows[0]["ttft_p50_us"] <= rows[0]["p50_us"]


def test_no_gpu_no_fallback(server):
    """Without the injected test double the tool needs the device and says so."""
    import ctypes

    from client_b200 import _native

    n = ctypes.c_int(0)
    if _native.load().tb200_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(Exception):
        cli.main(["-m", "simple", "-u", server["http"], "--shared-memory", "none", "--concurrency-range", "1", "-p", "100", "-r", "1"])


@pytest.mark.parametrize("protocol", ["http", "grpc"])
def test_string_inputs_over_the_wire(server, protocol):
    """BYTES inputs of fixed-length strings (--string-length) through the wire path."""
    rows = cli.main(["-m", "string_identity", "-u", server[protocol], "-i", protocol, "--shared-memory", "none",
                     "--string-length", "24", "--concurrency-range", "2", "-p", "200", "-r", "3", "--json"],
                    staging_factory=HostStaging)
    assert rows[0]["count"] > 3 and rows[0]["failed"] == 0 and rows[0]["input_bytes"] == 8 * 28

