"""HTTP response model.

Drop-in for ``tritonclient.http.InferResult`` (reference:
src/python/library/tritonclient/http/_infer_result.py:41-242): JSON header +
binary tensors indexed by ``binary_data_size``; gzip / deflate bodies.
"""

import gzip
import json
import zlib

import numpy as np

from ..utils import (
    deserialize_bf16_tensor,
    deserialize_bytes_tensor,
    raise_error,
    triton_to_np_dtype,
)


class _BufferResponse:
    """Minimal response interface (``get`` / ``read``) over bytes in memory."""

    def __init__(self, body, headers=None):
        self._body = body
        self._pos = 0
        self._headers = headers or {}

    def get(self, key):
        return self._headers.get(key)

    def read(self, length=-1):
        if length == -1:
            chunk = self._body[self._pos:]
            self._pos = len(self._body)
        else:
            chunk = self._body[self._pos:self._pos + length]
            self._pos += length
        return chunk


class InferResult:
    """Holds the response of an inference request.

    Parameters
    ----------
    response : object with ``get(header)`` and ``read(length=-1)``
        The inference response from the server.
    verbose : bool
        If True print the JSON header.
    """

    def __init__(self, response, verbose):
        header_length = response.get("Inference-Header-Content-Length")
        encoding = response.get("Content-Encoding")
        if encoding == "gzip":
            response = _BufferResponse(gzip.decompress(response.read()))
        elif encoding == "deflate":
            response = _BufferResponse(zlib.decompress(response.read()))

        self._buffer = b""
        self._output_name_to_buffer_map = {}
        if header_length is None:
            content = response.read()
            if verbose:
                print(content)
            try:
                self._result = json.loads(content)
            except UnicodeDecodeError as e:
                raise_error(
                    f"Failed to encode using UTF-8. Please use binary_data=True, if"
                    f" you want to pass a byte array. UnicodeError: {e}"
                )
            return
        content = response.read(length=int(header_length))
        if verbose:
            print(content)
        self._result = json.loads(content)
        self._buffer = response.read()
        cursor = 0
        for output in self._result["outputs"]:
            size = (output.get("parameters") or {}).get("binary_data_size")
            if size is not None:
                self._output_name_to_buffer_map[output["name"]] = cursor
                cursor += size

    @classmethod
    def from_response_body(cls, response_body, verbose=False, header_length=None, content_encoding=None):
        """Build an InferResult from raw response bytes (reference :108-155)."""
        headers = {"Inference-Header-Content-Length": header_length, "Content-Encoding": content_encoding}
        return cls(_BufferResponse(response_body, headers), verbose)

    def as_numpy(self, name):
        """The named output as a numpy array, or None when absent (reference :157-210)."""
        for output in self._result.get("outputs") or []:
            if output["name"] != name:
                continue
            datatype = output["datatype"]
            size = (output.get("parameters") or {}).get("binary_data_size")
            if size is None:
                array = np.array(output["data"], dtype=triton_to_np_dtype(datatype))
            elif size == 0:
                array = np.empty(0)
            else:
                start = self._output_name_to_buffer_map[name]
                chunk = self._buffer[start:start + size]
                if datatype == "BYTES":
                    array = deserialize_bytes_tensor(chunk)
                elif datatype == "BF16":
                    array = deserialize_bf16_tensor(chunk)
                else:
                    array = np.frombuffer(chunk, dtype=triton_to_np_dtype(datatype))
            return array.reshape(output["shape"])
        return None

    def get_output(self, name):
        """The JSON dict of the named output, or None."""
        for output in self._result["outputs"]:
            if output["name"] == name:
                return output
        return None

    def get_response(self):
        """The complete response header as a dict."""
        return self._result
