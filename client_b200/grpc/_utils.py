"""gRPC helpers: error mapping and request assembly.

Drop-in for ``tritonclient.grpc._utils`` (reference:
src/python/library/tritonclient/grpc/_utils.py:33-154).
"""

import grpc

from ..utils import (
    TRITON_RESERVED_REQUEST_PARAMS,
    TRITON_RESERVED_REQUEST_PARAMS_PREFIX,
    InferenceServerException,
    raise_error,
)
from . import service_pb2


def get_error_grpc(rpc_error):
    """grpc.RpcError -> InferenceServerException (reference :33-50)."""
    return InferenceServerException(
        msg=rpc_error.details(),
        status=str(rpc_error.code()),
        debug_details=rpc_error.debug_error_string(),
    )


def get_cancelled_error(msg=None):
    """InferenceServerException for a locally cancelled RPC (reference :53-62)."""
    if not msg:
        msg = "Locally cancelled by application!"
    return InferenceServerException(msg=msg, status="StatusCode.CANCELLED")


def raise_error_grpc(rpc_error):
    """Raise the InferenceServerException of a grpc.RpcError (reference :65-78)."""
    raise get_error_grpc(rpc_error) from None


def _set_parameter(slot, key, value):
    """Store a custom request parameter with the reference's type dispatch
    (str, then bool before int, then float; :122-137)."""
    if isinstance(value, str):
        slot[key].string_param = value
    elif isinstance(value, bool):
        slot[key].bool_param = value
    elif isinstance(value, int):
        slot[key].int64_param = value
    elif isinstance(value, float):
        slot[key].double_param = value
    else:
        raise_error(f'The parameter datatype "{type(value)}" for key "{key}" is not supported.')


def _get_inference_request(model_name, inputs, model_version, request_id, outputs, sequence_id,
                           sequence_start, sequence_end, priority, timeout, parameters):
    """Build the ``ModelInferRequest``: inputs in order, one ``raw_input_contents``
    entry per input that carries data (reference :80-139)."""
    request = service_pb2.ModelInferRequest(model_name=model_name, model_version=model_version)
    if request_id != "":
        request.id = request_id
    for entry in inputs:
        request.inputs.append(entry._get_tensor())
        content = entry._get_content()
        if content is not None:
            request.raw_input_contents.append(content)
    for entry in outputs or ():
        request.outputs.append(entry._get_tensor())
    if sequence_id != 0 and sequence_id != "":
        if isinstance(sequence_id, str):
            request.parameters["sequence_id"].string_param = sequence_id
        else:
            request.parameters["sequence_id"].int64_param = sequence_id
        request.parameters["sequence_start"].bool_param = sequence_start
        request.parameters["sequence_end"].bool_param = sequence_end
    if priority != 0:
        request.parameters["priority"].uint64_param = priority
    if timeout is not None:
        request.parameters["timeout"].int64_param = timeout
    for key, value in (parameters or {}).items():
        if key in TRITON_RESERVED_REQUEST_PARAMS or key.startswith(TRITON_RESERVED_REQUEST_PARAMS_PREFIX):
            raise_error(f'Parameter "{key}" is a reserved parameter and cannot be specified.')
        _set_parameter(request.parameters, key, value)
    return request


def _grpc_compression_type(algorithm_str):
    """'deflate' / 'gzip' / None -> grpc.Compression (reference :142-154)."""
    if algorithm_str is None:
        return grpc.Compression.NoCompression
    lowered = algorithm_str.lower()
    if lowered == "deflate":
        return grpc.Compression.Deflate
    if lowered == "gzip":
        return grpc.Compression.Gzip
    print(
        "The provided client-side compression algorithm is not supported... "
        "using no compression"
    )
    return grpc.Compression.NoCompression
