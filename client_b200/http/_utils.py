"""HTTP wire helpers: request body assembly, error mapping, query strings.

Drop-in for ``tritonclient.http._utils`` (reference:
src/python/library/tritonclient/http/_utils.py:40-151).  The JSON header is
emitted compactly in insertion order, which is what python-rapidjson produces
for the reference; ``json_size`` is the character count of the header like
there.
"""

import json
from urllib.parse import quote_plus

from ..utils import (
    TRITON_RESERVED_REQUEST_PARAMS,
    TRITON_RESERVED_REQUEST_PARAMS_PREFIX,
    InferenceServerException,
    raise_error,
)


def _dumps(obj):
    return json.dumps(obj, separators=(",", ":"))


def _get_error(response):
    """InferenceServerException for a non-200 response, else None (reference :40-65)."""
    if response.status_code == 200:
        return None
    body = None
    try:
        body = response.read().decode("utf-8")
        payload = (
            json.loads(body)
            if len(body)
            else {"error": "client received an empty response from the server."}
        )
        return InferenceServerException(msg=payload["error"], status=str(response.status_code))
    except Exception as e:
        return InferenceServerException(
            msg=f"an exception occurred in the client while decoding the response: {e}\nresponse: {body}",
            status=str(response.status_code),
            debug_details=body,
        )


def _raise_if_error(response):
    """Raise for a non-success response (reference :68-75)."""
    error = _get_error(response)
    if error is not None:
        raise error


def _get_query_string(query_params):
    """``k=v&k=v`` with quote_plus; list values repeat the key (reference :78-87)."""
    pairs = []
    for key, value in query_params.items():
        values = value if isinstance(value, list) else [value]
        for item in values:
            pairs.append("%s=%s" % (quote_plus(key), quote_plus(str(item))))
    return "&".join(pairs)


def _request_header_dict(inputs, request_id, outputs, sequence_id, sequence_start, sequence_end,
                         priority, timeout, custom_parameters):
    """The JSON object of an inference request; insertion order = wire order
    (SURVEY.md section 9.1)."""
    parameters = {}
    request = {}
    if request_id != "":
        request["id"] = request_id
    if sequence_id != 0 and sequence_id != "":
        parameters["sequence_id"] = sequence_id
        parameters["sequence_start"] = sequence_start
        parameters["sequence_end"] = sequence_end
    if priority != 0:
        parameters["priority"] = priority
    if timeout is not None:
        parameters["timeout"] = timeout
    request["inputs"] = [entry._get_tensor() for entry in inputs]
    if outputs:
        request["outputs"] = [entry._get_tensor() for entry in outputs]
    else:
        # nothing requested explicitly: ask for every output in binary form
        parameters["binary_data_output"] = True
    for key, value in (custom_parameters or {}).items():
        if key in TRITON_RESERVED_REQUEST_PARAMS or key.startswith(TRITON_RESERVED_REQUEST_PARAMS_PREFIX):
            raise_error(f'Parameter "{key}" is a reserved parameter and cannot be specified.')
        parameters[key] = value
    if parameters:
        request["parameters"] = parameters
    return request


def _get_inference_request(inputs, request_id, outputs, sequence_id, sequence_start, sequence_end,
                           priority, timeout, custom_parameters):
    """(body, json_size): JSON header followed by the raw tensors in input order;
    json_size is None when the body is JSON only (reference :90-151)."""
    header = _dumps(
        _request_header_dict(inputs, request_id, outputs, sequence_id, sequence_start, sequence_end,
                             priority, timeout, custom_parameters)
    )
    encoded = header.encode()
    blobs = [entry._get_binary_data() for entry in inputs]
    blobs = [b for b in blobs if b is not None]
    if not blobs:
        return encoded, None
    return b"".join([encoded] + blobs), len(header)
