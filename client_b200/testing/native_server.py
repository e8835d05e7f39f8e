"""Launcher for the native KServe-v2 stand-in server (csrc/mock_server.cu).

    python -m client_b200.testing.native_server --port 8000 --device 0

Serves the CUDA-shared-memory subset of the HTTP protocol for models ``densenet_onnx``
and ``simple`` with the model run as a CUDA kernel on the client's IPC-mapped regions, and
-- with ``--grpc-port`` -- inference.GRPCInferenceService over its own HTTP/2 core: the same
models plus ``bert_large`` and the decoupled ``llama3_8b`` with tensors in the messages
(ModelInfer and ModelStreamInfer).  It has to be its own process (a CUDA IPC handle cannot be
opened by the process that exported it).  Tooling for loopback load runs; tests that need the
wider protocol (system shm, BYTES, sequences) use ``client_b200.testing.mock_server``.
"""

import argparse
import ctypes
import signal
import sys
import time

from .. import _native


class NativeServer:
    def __init__(self, host="127.0.0.1", port=0, device=0, grpc_port=None):
        self._lib = _native.load()
        p = ctypes.c_int(port)
        h = ctypes.c_void_p()
        if grpc_port is None:
            _native.check(self._lib.tb200_mock_server_start(host.encode(), ctypes.byref(p), device, ctypes.byref(h)))
            self.grpc_port = None
        else:
            g = ctypes.c_int(grpc_port)
            _native.check(self._lib.tb200_mock_server_start2(host.encode(), ctypes.byref(p), ctypes.byref(g), device, ctypes.byref(h)))
            self.grpc_port = g.value
        self._h, self.host, self.port = h, host, p.value

    @property
    def url(self):
        return "%s:%d" % (self.host, self.port)

    @property
    def requests(self):
        return int(self._lib.tb200_mock_server_requests(self._h)) if self._h else 0

    @property
    def batches(self):
        return int(self._lib.tb200_mock_server_batches(self._h)) if self._h else 0

    def stop(self):
        if getattr(self, "_h", None):
            self._lib.tb200_mock_server_stop(self._h)
            self._h = None


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--grpc-port", type=int, default=None, help="also serve gRPC on this port (0 = pick one)")
    ap.add_argument("--pin-cpus", action="store_true", help="pin the server's threads to its GPU's share of the local NUMA node (perf/topology.py)")
    args = ap.parse_args(argv)
    if args.pin_cpus:
        from ..perf.topology import pin_for

        pin_for(args.device, "server")  # before the server's threads exist
    srv = NativeServer(args.host, args.port, args.device, args.grpc_port)
    print("native mock server listening on %s%s" % (srv.url, "" if srv.grpc_port is None else " grpc=%s:%d" % (srv.host, srv.grpc_port)), flush=True)
    stop = []
    signal.signal(signal.SIGTERM, lambda *a: stop.append(1))
    signal.signal(signal.SIGINT, lambda *a: stop.append(1))
    while not stop:
        time.sleep(0.2)
    n, b = srv.requests, srv.batches
    srv.stop()
    print("served %d inference requests in %d model launches" % (n, b), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
