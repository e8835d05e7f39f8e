"""Datatype maps and tensor (de)serialisation helpers.

Drop-in for ``tritonclient.utils`` (reference:
src/python/library/tritonclient/utils/__init__.py).  Same names, arguments,
return types and error texts; the BYTES / BF16 codecs are vectorised instead of
the reference's per-element Python loops (:244-257, :327-331, :358-363).
"""

import numpy as np

from ._shared_memory_tensor import SharedMemoryTensor  # noqa: F401  (re-export, ref :33)

# Request parameters the server reserves (reference :39-48).
TRITON_RESERVED_REQUEST_PARAMS = [
    "sequence_id",
    "sequence_start",
    "sequence_end",
    "priority",
    "timeout",
    "headers",
    "binary_data_output",
]
TRITON_RESERVED_REQUEST_PARAMS_PREFIX = "triton_"


class InferenceServerException(Exception):
    """Non-success status from the client or the server (reference :86-145).

    Parameters
    ----------
    msg : str
        Brief description of the error.
    status : str
        Error code, if any.
    debug_details : str
        Additional details, if any.
    """

    def __init__(self, msg, status=None, debug_details=None):
        self._msg = msg
        self._status = status
        self._debug_details = debug_details

    def __str__(self):
        text = super().__str__() if self._msg is None else self._msg
        if self._status is not None:
            text = "[" + self._status + "] " + text
        return text

    def message(self):
        """The message of this exception, or None."""
        return self._msg

    def status(self):
        """The status (error code) of this exception, or None."""
        return self._status

    def debug_details(self):
        """Detailed information for debugging, or None."""
        return self._debug_details


def raise_error(msg):
    """Raise :py:class:`InferenceServerException` with ``msg`` (reference :51-55)."""
    raise InferenceServerException(msg=msg) from None


# (triton name, numpy type) in the order of the reference's if-chains (:148-205).
_DTYPE_TABLE = (
    ("BOOL", bool),
    ("INT8", np.int8),
    ("INT16", np.int16),
    ("INT32", np.int32),
    ("INT64", np.int64),
    ("UINT8", np.uint8),
    ("UINT16", np.uint16),
    ("UINT32", np.uint32),
    ("UINT64", np.uint64),
    ("FP16", np.float16),
    ("FP32", np.float32),
    ("FP64", np.float64),
)
_TRITON_TO_NP = dict(_DTYPE_TABLE)
_TRITON_TO_NP["BF16"] = np.float32  # numpy has no bfloat16 (reference :199)
_TRITON_TO_NP["BYTES"] = np.object_
_NP_TO_TRITON = {np.dtype(npt): name for name, npt in _DTYPE_TABLE}


def np_to_triton_dtype(np_dtype):
    """numpy dtype -> Triton datatype name, None when unmapped (reference :148-175)."""
    try:
        dt = np.dtype(np_dtype)
    except TypeError:
        return None
    name = _NP_TO_TRITON.get(dt)
    if name is not None:
        return name
    if dt == np.object_ or dt.type == np.bytes_:
        return "BYTES"
    return None


def triton_to_np_dtype(dtype):
    """Triton datatype name -> numpy type, None when unknown (reference :178-205)."""
    return _TRITON_TO_NP.get(dtype)


def _as_object_scalar(payload):
    """0-d object ndarray holding ``payload`` -- the container the reference
    returns from its serialisers (:258-261, :332-335)."""
    boxed = np.empty((), dtype=np.object_)
    boxed[()] = payload
    return boxed


def _element_bytes(input_tensor):
    """Row-major list of the byte strings of a BYTES tensor (reference :244-254)."""
    if input_tensor.dtype == np.object_:
        out = []
        for item in input_tensor.ravel(order="C").tolist():
            out.append(item if type(item) == bytes else str(item).encode("utf-8"))
        return out
    # np.bytes_ ('S') arrays: numpy strips trailing NULs exactly like .item()
    return input_tensor.ravel(order="C").tolist()


def serialized_byte_size(tensor_value):
    """Total payload bytes of an object tensor (reference :58-83)."""
    if tensor_value.dtype != np.object_:
        raise_error("The tensor_value dtype must be np.object_")
    if tensor_value.size == 0:
        return 0
    return sum(len(item) for item in tensor_value.ravel(order="C").tolist())


def serialize_byte_tensor(input_tensor):
    """BYTES tensor -> ``<u32 little-endian length><payload>`` per element,
    row-major, boxed in a 0-d object array (reference :208-261)."""
    if input_tensor.size == 0:
        return np.empty([0], dtype=np.object_)
    if (input_tensor.dtype != np.object_) and (input_tensor.dtype.type != np.bytes_):
        raise_error("cannot serialize bytes tensor: invalid datatype")
    items = _element_bytes(input_tensor)
    lengths = np.fromiter((len(s) for s in items), dtype=np.int64, count=len(items))
    if lengths.size and int(lengths.max()) > 0xFFFFFFFF:
        raise_error("cannot serialize bytes tensor: element larger than 4 GiB")
    total = int(lengths.sum()) + 4 * len(items)
    frame = bytearray(total)
    prefixes = lengths.astype("<u4").tobytes()
    pos = 0
    for i, s in enumerate(items):
        frame[pos : pos + 4] = prefixes[4 * i : 4 * i + 4]
        pos += 4
        n = len(s)
        frame[pos : pos + n] = s
        pos += n
    return _as_object_scalar(bytes(frame))


def deserialize_bytes_tensor(encoded_tensor):
    """Inverse of :py:func:`serialize_byte_tensor`: 1-D object array of bytes
    (reference :264-291)."""
    view = memoryview(encoded_tensor).cast("B") if not isinstance(encoded_tensor, (bytes, bytearray)) else encoded_tensor
    strs = []
    offset = 0
    end = len(view)
    while offset < end:
        if offset + 4 > end:
            raise_error("malformed BYTES tensor: truncated length prefix")
        n = int.from_bytes(view[offset : offset + 4], "little")
        offset += 4
        if offset + n > end:
            raise_error("malformed BYTES tensor: truncated element")
        strs.append(bytes(view[offset : offset + n]))
        offset += n
    return np.array(strs, dtype=np.object_)


def serialize_bf16_tensor(input_tensor):
    """float32 tensor -> bfloat16 wire bytes by truncation (upper 16 bits of
    each fp32, no rounding), boxed like the reference (:294-335)."""
    if input_tensor.size == 0:
        return np.empty([0], dtype=np.object_)
    if input_tensor.dtype != np.float32:
        raise_error("cannot serialize bf16 tensor: invalid datatype")
    bits = np.ascontiguousarray(input_tensor, dtype="<f4").reshape(-1).view("<u4")
    return _as_object_scalar((bits >> 16).astype("<u2").tobytes())


def deserialize_bf16_tensor(encoded_tensor):
    """bfloat16 wire bytes -> float32 array.  Shape (N, 1) like the reference,
    which appends 1-tuples (:355-363); callers reshape."""
    raw = np.frombuffer(encoded_tensor, dtype="<u2")
    if raw.size == 0:
        return np.array([], dtype=np.float32)
    return (raw.astype("<u4") << 16).view("<f4").astype(np.float32).reshape(-1, 1)
