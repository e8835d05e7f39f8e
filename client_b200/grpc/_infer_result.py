"""gRPC response model.

Drop-in for ``tritonclient.grpc.InferResult`` (reference:
src/python/library/tritonclient/grpc/_infer_result.py:34-158).
"""

import json

import numpy as np
from google.protobuf.json_format import MessageToJson

from ..utils import deserialize_bf16_tensor, deserialize_bytes_tensor, triton_to_np_dtype


class InferResult:
    """Holds a ``ModelInferResponse``.

    Parameters
    ----------
    result : protobuf message
        The ModelInferResponse returned by the server.
    """

    def __init__(self, result):
        self._result = result

    def as_numpy(self, name):
        """The named output as a numpy array, or None when absent.  Output i reads
        ``raw_output_contents[i]`` (reference :47-96)."""
        raws = self._result.raw_output_contents
        for index, output in enumerate(self._result.outputs):
            if output.name != name:
                continue
            shape = list(output.shape)
            datatype = output.datatype
            if index < len(raws):
                if datatype == "BYTES":
                    array = deserialize_bytes_tensor(raws[index])
                elif datatype == "BF16":
                    array = deserialize_bf16_tensor(raws[index])
                else:
                    array = np.frombuffer(raws[index], dtype=triton_to_np_dtype(datatype))
            elif len(output.contents.bytes_contents) != 0:
                array = np.array(list(output.contents.bytes_contents), dtype=np.object_)
            else:
                array = np.empty(0)
            return array.reshape(shape)
        return None

    def get_output(self, name, as_json=False):
        """The named InferOutputTensor (message, or dict with ``as_json``), or None."""
        for output in self._result.outputs:
            if output.name == name:
                if as_json:
                    return json.loads(MessageToJson(output, preserving_proto_field_name=True))
                return output
        return None

    def get_response(self, as_json=False):
        """The complete ModelInferResponse (message, or dict with ``as_json``)."""
        if as_json:
            return json.loads(MessageToJson(self._result, preserving_proto_field_name=True))
        return self._result
