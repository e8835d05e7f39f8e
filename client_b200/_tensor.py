"""Tensor checks and wire encoding shared by the HTTP and gRPC request models.

Restates the validation of the reference's ``InferInput.set_data_from_numpy``
(PY/http/_infer_input.py:128-160, PY/grpc/_infer_input.py:125-155; PY =
src/python/library/tritonclient) once instead of twice: same checks, same
message texts.
"""

import numpy as np

from .utils import (
    np_to_triton_dtype,
    raise_error,
    serialize_bf16_tensor,
    serialize_byte_tensor,
    triton_to_np_dtype,
)

SHM_KEYS = ("shared_memory_region", "shared_memory_byte_size", "shared_memory_offset")


def check_numpy_input(datatype, shape, input_tensor):
    """Raise InferenceServerException unless ``input_tensor`` is an ndarray whose
    dtype and shape match the declared ``datatype`` / ``shape``."""
    if not isinstance(input_tensor, (np.ndarray,)):
        raise_error("input_tensor must be a numpy array")
    if datatype == "BF16":
        # numpy has no bfloat16: BF16 inputs are supplied as float32 (DLIS-3986)
        expected = triton_to_np_dtype(datatype)
        if input_tensor.dtype != expected:
            raise_error(
                "got unexpected datatype {} from numpy array, expected {} for BF16 type".format(
                    input_tensor.dtype, expected
                )
            )
    else:
        actual = np_to_triton_dtype(input_tensor.dtype)
        if datatype != actual:
            raise_error(
                "got unexpected datatype {} from numpy array, expected {}".format(actual, datatype)
            )
    declared = list(shape)
    ok = len(declared) == input_tensor.ndim and all(
        declared[i] == input_tensor.shape[i] for i in range(len(declared))
    )
    if not ok:
        raise_error(
            "got unexpected numpy array shape [{}], expected [{}]".format(
                str(input_tensor.shape)[1:-1], str(declared)[1:-1]
            )
        )


def wire_bytes(datatype, input_tensor):
    """Row-major little-endian wire bytes of the tensor (a copy: the caller may
    mutate the array afterwards, like with the reference's ``tobytes()``)."""
    if datatype == "BYTES":
        boxed = serialize_byte_tensor(input_tensor)
        return boxed.item() if boxed.size > 0 else b""
    if datatype == "BF16":
        boxed = serialize_bf16_tensor(input_tensor)
        return boxed.item() if boxed.size > 0 else b""
    return input_tensor.tobytes()
