"""Opt-in native transport for ``InferenceServerClient.infer`` (gRPC).

The reference's ``infer`` goes through grpcio's C core and its Python glue
(src/python/library/tritonclient/grpc/_client.py:1445-1572).  With
``InferenceServerClient(url, transport="native")`` (or ``TB200_GRPC_TRANSPORT=native``) the
serialised ``ModelInferRequest`` is handed to libtb200client's own HTTP/2 channel
(client_b200/cpp/tb200_grpc_channel.h) instead: one blocking call with the GIL released, the
response bytes come back and are parsed into the same ``ModelInferResponse``.  Cleartext
connections only; every other RPC and the streams stay on grpcio.
"""

import ctypes
import os

from ..utils import InferenceServerException

_STATUS = ["OK", "CANCELLED", "UNKNOWN", "INVALID_ARGUMENT", "DEADLINE_EXCEEDED", "NOT_FOUND", "ALREADY_EXISTS",
           "PERMISSION_DENIED", "RESOURCE_EXHAUSTED", "FAILED_PRECONDITION", "ABORTED", "OUT_OF_RANGE", "UNIMPLEMENTED",
           "INTERNAL", "UNAVAILABLE", "DATA_LOSS", "UNAUTHENTICATED"]
_lib = None


def _load():
    global _lib
    if _lib is None:
        from .. import _native

        _native.load()  # libtb200.so first: libtb200client.so links against it
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libtb200client.so")
        if not os.path.exists(path):
            from ..build import build_cpp_client

            build_cpp_client()
        lib = ctypes.CDLL(path)
        lib.tb200c_grpc_channel_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        lib.tb200c_grpc_channel_close.argtypes = [ctypes.c_void_p]
        lib.tb200c_grpc_channel_close.restype = None
        lib.tb200c_grpc_unary.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_char_p),
                                          ctypes.c_int, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64),
                                          ctypes.c_char_p, ctypes.c_uint64]
        lib.tb200c_free.argtypes = [ctypes.c_void_p]
        lib.tb200c_free.restype = None
        _lib = lib
    return _lib


class NativeChannel:
    """One HTTP/2 connection to ``url`` (re-made after a failure); ``unary`` is thread safe."""

    def __init__(self, url):
        self._lib = _load()
        h = ctypes.c_void_p()
        if self._lib.tb200c_grpc_channel_open(url.encode(), ctypes.byref(h)) != 0:
            raise InferenceServerException(msg="cannot create the native gRPC channel")
        self._h = h
        self._message = ctypes.create_string_buffer(4096)

    def unary(self, path, request_bytes, metadata=(), timeout=None):
        """Serialised request -> serialised response; raises InferenceServerException with the
        status / message grpcio would report."""
        pairs = []
        for k, v in metadata:
            pairs += [str(k).lower().encode(), str(v).encode()]
        arr = (ctypes.c_char_p * max(len(pairs), 1))(*pairs)
        resp = ctypes.c_void_p()
        n = ctypes.c_uint64()
        message = ctypes.create_string_buffer(4096)  # per call: several threads may share the channel
        timeout_us = 0 if timeout is None else max(1, int(float(timeout) * 1e6))
        status = self._lib.tb200c_grpc_unary(self._h, path.encode(), request_bytes, len(request_bytes), arr, len(pairs) // 2, timeout_us,
                                             ctypes.byref(resp), ctypes.byref(n), message, len(message))
        if status != 0:
            name = _STATUS[status] if 0 <= status < len(_STATUS) else "UNKNOWN"
            raise InferenceServerException(msg=message.value.decode("utf-8", "replace"), status="StatusCode." + name)
        try:
            return ctypes.string_at(resp, n.value)
        finally:
            self._lib.tb200c_free(resp)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.tb200c_grpc_channel_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
