"""gRPC request model: one input tensor.

Drop-in for ``tritonclient.grpc.InferInput`` (reference:
src/python/library/tritonclient/grpc/_infer_input.py:36-219): the tensor is held
as an ``InferInputTensor`` message plus its raw bytes for ``raw_input_contents``.
"""

from .._tensor import SHM_KEYS, check_numpy_input, wire_bytes
from . import service_pb2


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
        self._input = service_pb2.ModelInferRequest.InferInputTensor(name=name, datatype=datatype)
        self._input.shape.extend(shape)
        self._raw_content = None

    def name(self):
        """The name of the input."""
        return self._input.name

    def datatype(self):
        """The datatype of the input."""
        return self._input.datatype

    def shape(self):
        """The shape of the input."""
        return self._input.shape

    def set_shape(self, shape):
        """Set the shape; returns the updated input."""
        del self._input.shape[:]
        self._input.shape.extend(shape)
        return self

    def set_data_from_numpy(self, input_tensor):
        """Take the tensor data from a numpy array (raises
        InferenceServerException on a dtype / shape mismatch); returns the updated
        input."""
        check_numpy_input(self._input.datatype, self._input.shape, input_tensor)
        for key in SHM_KEYS:
            self._input.parameters.pop(key, None)
        self._raw_content = wire_bytes(self._input.datatype, input_tensor)
        return self

    def set_shared_memory(self, region_name, byte_size, offset=0):
        """Take the tensor data from a registered shared memory region; returns the
        updated input."""
        self._input.ClearField("contents")
        self._raw_content = None
        self._input.parameters["shared_memory_region"].string_param = region_name
        self._input.parameters["shared_memory_byte_size"].int64_param = byte_size
        if offset != 0:
            self._input.parameters["shared_memory_offset"].int64_param = offset
        return self

    def _get_tensor(self):
        """The underlying InferInputTensor message."""
        return self._input

    def _get_content(self):
        """Raw tensor bytes for ``raw_input_contents`` (None for shared memory)."""
        return self._raw_content
