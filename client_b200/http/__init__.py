"""HTTP/REST client (drop-in for ``tritonclient.http``; reference:
src/python/library/tritonclient/http/__init__.py:29-53)."""

from ..utils import *  # noqa: F401,F403
from .._plugin import InferenceServerClientPlugin
from .._request import Request
from ._client import InferAsyncRequest, InferenceServerClient, set_device_compression
from ._infer_input import InferInput
from ._infer_result import InferResult
from ._requested_output import InferRequestedOutput
from ._utils import InferenceServerException

__all__ = [
    "InferenceServerClientPlugin",
    "Request",
    "InferenceServerClient",
    "InferInput",
    "InferRequestedOutput",
    "InferResult",
    "InferAsyncRequest",
    "InferenceServerException",
    "set_device_compression",
]
