"""gRPC request model: one requested output.

Drop-in for ``tritonclient.grpc.InferRequestedOutput`` (reference:
src/python/library/tritonclient/grpc/_requested_output.py:33-108).
"""

from ..utils import raise_error
from . import service_pb2


class InferRequestedOutput:
    """Describes a requested output tensor.

    Parameters
    ----------
    name : str
        The name of the output tensor.
    class_count : int
        Number of classifications to request; 0 (default) requests the tensor.
    """

    def __init__(self, name, class_count=0):
        self._output = service_pb2.ModelInferRequest.InferRequestedOutputTensor(name=name)
        if class_count != 0:
            self._output.parameters["classification"].int64_param = class_count

    def name(self):
        """The name of the output."""
        return self._output.name

    def set_shared_memory(self, region_name, byte_size, offset=0):
        """Have the server write this output into a registered shared memory region."""
        if "classification" in self._output.parameters:
            raise_error("shared memory can't be set on classification output")
        self._output.parameters["shared_memory_region"].string_param = region_name
        self._output.parameters["shared_memory_byte_size"].int64_param = byte_size
        if offset != 0:
            self._output.parameters["shared_memory_offset"].int64_param = offset

    def unset_shared_memory(self):
        """Undo :py:meth:`set_shared_memory`."""
        for key in ("shared_memory_region", "shared_memory_byte_size", "shared_memory_offset"):
            self._output.parameters.pop(key, None)

    def _get_tensor(self):
        """The underlying InferRequestedOutputTensor message."""
        return self._output
