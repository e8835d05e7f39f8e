"""gRPC client (drop-in for ``tritonclient.grpc``; reference:
src/python/library/tritonclient/grpc/__init__.py:29-73).  ``service_pb2``,
``service_pb2_grpc`` and ``model_config_pb2`` are built at import time from the
schema table in ``_proto.py`` (no protoc in this image)."""

import sys as _sys

import grpc  # noqa: F401

from ._proto import build_modules as _build_modules

service_pb2, service_pb2_grpc, model_config_pb2 = _build_modules()
for _m in (service_pb2, service_pb2_grpc, model_config_pb2):
    _sys.modules[_m.__name__] = _m

from ..utils import *  # noqa: E402,F401,F403
from .._plugin import InferenceServerClientPlugin  # noqa: E402
from .._request import Request  # noqa: E402
from ._client import MAX_GRPC_MESSAGE_SIZE, InferenceServerClient, KeepAliveOptions  # noqa: E402
from ._infer_input import InferInput  # noqa: E402
from ._infer_result import InferResult  # noqa: E402
from ._requested_output import InferRequestedOutput  # noqa: E402
from ._utils import raise_error, raise_error_grpc  # noqa: E402

__all__ = [
    "InferenceServerClientPlugin",
    "Request",
    "InferenceServerClient",
    "InferInput",
    "InferRequestedOutput",
    "InferResult",
    "KeepAliveOptions",
    "InferenceServerException",
]
