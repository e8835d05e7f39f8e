"""HTTP request model: one input tensor.

Drop-in for ``tritonclient.http.InferInput`` (reference:
src/python/library/tritonclient/http/_infer_input.py:38-272): same constructor,
setters returning ``self``, error texts, and the same JSON dict from
``_get_tensor()``.
"""

import numpy as np

from .._tensor import SHM_KEYS, check_numpy_input, wire_bytes
from ..utils import raise_error


def _json_strings(input_tensor):
    """BYTES elements as JSON strings (binary_data=False), reference :166-188."""
    out = []
    item = None
    try:
        for item in input_tensor.ravel(order="C").tolist():
            if input_tensor.dtype == np.object_ and type(item) != bytes:
                out.append(str(item))
            else:
                out.append(str(item, encoding="utf-8"))
    except UnicodeDecodeError:
        raise_error(
            f'Failed to encode "{item}" using UTF-8. Please use binary_data=True, if'
            " you want to pass a byte array."
        )
    return out


class InferInput:
    """Describes one input tensor of an inference request.

    Parameters
    ----------
    name : str
        The name of the input.
    shape : list
        The shape of the input.
    datatype : str
        The Triton datatype of the input.
    """

    def __init__(self, name, shape, datatype):
        self._name = name
        self._shape = shape
        self._datatype = datatype
        self._parameters = {}
        self._data = None
        self._raw_data = None

    def name(self):
        """The name of the input."""
        return self._name

    def datatype(self):
        """The datatype of the input."""
        return self._datatype

    def shape(self):
        """The shape of the input."""
        return self._shape

    def set_shape(self, shape):
        """Set the shape; returns the updated input."""
        self._shape = shape
        return self

    def _drop_shm(self):
        for key in SHM_KEYS:
            self._parameters.pop(key, None)

    def set_data_from_numpy(self, input_tensor, binary_data=True):
        """Take the tensor data from a numpy array.

        ``binary_data=True`` (default) sends the tensor as raw bytes after the JSON
        header, otherwise as a JSON list.  Raises InferenceServerException on a
        dtype / shape mismatch.  Returns the updated input.
        """
        check_numpy_input(self._datatype, self._shape, input_tensor)
        self._drop_shm()
        if binary_data:
            self._data = None
            self._raw_data = wire_bytes(self._datatype, input_tensor)
            self._parameters["binary_data_size"] = len(self._raw_data)
            return self
        self._parameters.pop("binary_data_size", None)
        self._raw_data = None
        if self._datatype == "BF16":
            raise_error(
                "BF16 inputs must be sent as binary data over HTTP. Please set binary_data=True"
            )
        if self._datatype == "BYTES":
            self._data = _json_strings(input_tensor) if input_tensor.size > 0 else []
        else:
            self._data = input_tensor.ravel(order="C").tolist()
        return self

    def set_shared_memory(self, region_name, byte_size, offset=0):
        """Take the tensor data from a registered shared memory region; returns the
        updated input."""
        self._data = None
        self._raw_data = None
        self._parameters.pop("binary_data_size", None)
        self._parameters["shared_memory_region"] = region_name
        self._parameters["shared_memory_byte_size"] = byte_size
        if offset != 0:
            self._parameters["shared_memory_offset"] = offset
        return self

    def _get_binary_data(self):
        """Raw tensor bytes, or None when the data is inline JSON / in shared memory."""
        return self._raw_data

    def _get_tensor(self):
        """The JSON dict of this input (key order is part of the wire contract)."""
        tensor = {"name": self._name, "shape": self._shape, "datatype": self._datatype}
        if self._parameters:
            tensor["parameters"] = self._parameters
        inline = self._parameters.get("shared_memory_region") is None and self._raw_data is None
        if inline and self._data is not None:
            tensor["data"] = self._data
        return tensor
