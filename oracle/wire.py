"""CPU restatement of the reference client's wire encoders. TEST INFRASTRUCTURE.

Restates, function by function (paths under /root/reference/src/python/library/tritonclient):
  * utils/__init__.py:208-261   serialize_byte_tensor   -> bytes_tensor_wire
  * utils/__init__.py:294-335   serialize_bf16_tensor   -> bf16_tensor_wire
  * http/_infer_input.py:106-214, 254-272                -> HttpInput
  * http/_requested_output.py:51-117                     -> HttpOutput
  * http/_utils.py:90-151       _get_inference_request   -> http_request_body
  * http/_infer_result.py:54-210                         -> http_parse_response
  * grpc/_infer_input.py, grpc/_utils.py:80-139          -> grpc_request_bytes (a
    hand-rolled protobuf encoder following grpc_service.proto:439-706 in
    src/rust/triton-client/proto, so it is independent of the protobuf runtime
    the product uses)
Pinned by tests/golden/wire_golden.json, generated from the reference itself by
oracle/gen_golden.py (shimmed import), and by the
reference's known-answer vectors.

The reference dumps JSON with python-rapidjson (absent here); for the integer /
string / bool headers of the hot path, compact stdlib json in insertion order is
byte-identical.  Float-in-JSON formatting is unpinned.
"""

import json
import struct

import numpy as np

RESERVED = ["sequence_id", "sequence_start", "sequence_end", "priority", "timeout", "headers", "binary_data_output"]

_NP2T = [(bool, "BOOL"), (np.int8, "INT8"), (np.int16, "INT16"), (np.int32, "INT32"), (np.int64, "INT64"),
         (np.uint8, "UINT8"), (np.uint16, "UINT16"), (np.uint32, "UINT32"), (np.uint64, "UINT64"),
         (np.float16, "FP16"), (np.float32, "FP32"), (np.float64, "FP64")]


def np_to_triton(dt):
    for npt, name in _NP2T:
        if dt == npt:
            return name
    if dt == np.object_ or dt.type == np.bytes_:
        return "BYTES"
    return None


def bytes_tensor_wire(arr):
    """utils/__init__.py:242-257 -- per element <I len + payload, row-major."""
    if arr.size == 0:
        return b""
    out = []
    for obj in np.nditer(arr, flags=["refs_ok"], order="C"):
        item = obj.item()
        if arr.dtype == np.object_:
            s = item if type(item) == bytes else str(item).encode("utf-8")
        else:
            s = item
        out.append(struct.pack("<I", len(s)))
        out.append(s)
    return b"".join(out)


def bf16_tensor_wire(arr):
    """utils/__init__.py:325-331 -- struct.pack('<f', x)[2:4] per element."""
    if arr.size == 0:
        return b""
    return b"".join(struct.pack("<f", x)[2:4] for x in np.nditer(arr, order="C"))


def tensor_wire_bytes(arr, datatype):
    if datatype == "BYTES":
        return bytes_tensor_wire(arr)
    if datatype == "BF16":
        return bf16_tensor_wire(arr)
    return arr.tobytes()


class HttpInput:
    """http/_infer_input.py InferInput state machine."""

    def __init__(self, name, shape, datatype):
        self.name, self.shape, self.datatype = name, shape, datatype
        self.parameters = {}
        self.data = None
        self.raw = None

    def set_data(self, arr, binary_data=True):
        for k in ("shared_memory_region", "shared_memory_byte_size", "shared_memory_offset"):
            self.parameters.pop(k, None)
        if not binary_data:
            self.parameters.pop("binary_data_size", None)
            self.raw = None
            if self.datatype == "BYTES":
                self.data = []
                if arr.size > 0:
                    for obj in np.nditer(arr, flags=["refs_ok"], order="C"):
                        item = obj.item()
                        if arr.dtype == np.object_ and type(item) != bytes:
                            self.data.append(str(item))
                        else:
                            self.data.append(str(item, encoding="utf-8"))
            else:
                self.data = [v.item() for v in arr.flatten()]
        else:
            self.data = None
            self.raw = tensor_wire_bytes(arr, self.datatype)
            self.parameters["binary_data_size"] = len(self.raw)
        return self

    def set_shm(self, region, byte_size, offset=0):
        self.data = None
        self.raw = None
        self.parameters.pop("binary_data_size", None)
        self.parameters["shared_memory_region"] = region
        self.parameters["shared_memory_byte_size"] = byte_size
        if offset != 0:
            self.parameters["shared_memory_offset"] = offset
        return self

    def tensor(self):
        t = {"name": self.name, "shape": self.shape, "datatype": self.datatype}
        if self.parameters:
            t["parameters"] = self.parameters
        if self.parameters.get("shared_memory_region") is None and self.raw is None:
            if self.data is not None:
                t["data"] = self.data
        return t


class HttpOutput:
    """http/_requested_output.py InferRequestedOutput."""

    def __init__(self, name, binary_data=True, class_count=0):
        self.name = name
        self.parameters = {}
        if class_count != 0:
            self.parameters["classification"] = class_count
        self.binary = binary_data
        self.parameters["binary_data"] = binary_data

    def set_shm(self, region, byte_size, offset=0):
        if self.binary:
            self.parameters["binary_data"] = False
        self.parameters["shared_memory_region"] = region
        self.parameters["shared_memory_byte_size"] = byte_size
        if offset != 0:
            self.parameters["shared_memory_offset"] = offset
        return self

    def tensor(self):
        t = {"name": self.name}
        if self.parameters:
            t["parameters"] = self.parameters
        return t


def http_request_body(inputs, outputs=None, request_id="", sequence_id=0, sequence_start=False,
                      sequence_end=False, priority=0, timeout=None, parameters=None):
    """http/_utils.py:90-151 -> (body bytes, json_size or None)."""
    req = {}
    params = {}
    if request_id != "":
        req["id"] = request_id
    if sequence_id != 0 and sequence_id != "":
        params["sequence_id"] = sequence_id
        params["sequence_start"] = sequence_start
        params["sequence_end"] = sequence_end
    if priority != 0:
        params["priority"] = priority
    if timeout is not None:
        params["timeout"] = timeout
    req["inputs"] = [i.tensor() for i in inputs]
    if outputs:
        req["outputs"] = [o.tensor() for o in outputs]
    else:
        params["binary_data_output"] = True
    if parameters:
        for k, v in parameters.items():
            if k in RESERVED or k.startswith("triton_"):
                raise ValueError('Parameter "%s" is a reserved parameter and cannot be specified.' % k)
            params[k] = v
    if params:
        req["parameters"] = params
    text = json.dumps(req, separators=(",", ":"))
    chunks = [text.encode()]
    for i in inputs:
        if i.raw is not None:
            chunks.append(i.raw)
    if len(chunks) == 1:
        return chunks[0], None
    return b"".join(chunks), len(text)


def http_parse_response(body, header_length=None):
    """http/_infer_result.py:54-210 -> {name: ndarray}."""
    t2np = {name: npt for npt, name in _NP2T}
    if header_length is None:
        result, buf = json.loads(body), b""
    else:
        result, buf = json.loads(body[:header_length]), body[header_length:]
    out = {}
    idx = 0
    for o in result.get("outputs", []):
        p = o.get("parameters") or {}
        size = p.get("binary_data_size")
        dt = o["datatype"]
        if size is not None:
            chunk = buf[idx:idx + size]
            idx += size
            if size == 0:
                arr = np.empty(0)
            elif dt == "BYTES":
                items, pos = [], 0
                while pos < len(chunk):
                    n = struct.unpack_from("<I", chunk, pos)[0]
                    pos += 4
                    items.append(chunk[pos:pos + n])
                    pos += n
                arr = np.array(items, dtype=np.object_)
            elif dt == "BF16":
                arr = np.array([struct.unpack("<f", b"\x00\x00" + chunk[i:i + 2]) for i in range(0, len(chunk), 2)], dtype=np.float32)
            else:
                arr = np.frombuffer(chunk, dtype=t2np[dt])
        else:
            arr = np.array(o["data"], dtype=np.float32 if dt == "BF16" else (np.object_ if dt == "BYTES" else t2np[dt]))
        out[o["name"]] = arr.reshape(o["shape"])
    return out


# ---- protobuf wire (grpc_service.proto ModelInferRequest, fields 1-7) ------------
def _varint(n):
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _infer_parameter(value, kind):
    """InferParameter oneof: 1 bool, 2 int64, 3 string, 4 double, 5 uint64."""
    if kind == "bool":
        return _varint((1 << 3) | 0) + _varint(1 if value else 0)
    if kind == "int64":
        return _varint((2 << 3) | 0) + _varint(value)
    if kind == "string":
        return _ld(3, value.encode("utf-8"))
    if kind == "double":
        return _varint((4 << 3) | 1) + struct.pack("<d", value)
    if kind == "uint64":
        return _varint((5 << 3) | 0) + _varint(value)
    raise ValueError(kind)


def _map_entry(field, key, param_payload):
    return _ld(field, _ld(1, key.encode("utf-8")) + _ld(2, param_payload))


class GrpcInput:
    """grpc/_infer_input.py InferInput: name, datatype, shape, shm params, raw bytes."""

    def __init__(self, name, shape, datatype):
        self.name, self.shape, self.datatype = name, list(shape), datatype
        self.params = []  # ordered (key, kind, value); at most one used in pinned cases
        self.raw = None

    def set_data(self, arr):
        self.params = [p for p in self.params if not p[0].startswith("shared_memory_")]
        self.raw = tensor_wire_bytes(arr, self.datatype)
        return self

    def set_shm(self, region, byte_size, offset=0):
        self.raw = None
        self.params = [("shared_memory_region", "string", region), ("shared_memory_byte_size", "int64", byte_size)]
        if offset != 0:
            self.params.append(("shared_memory_offset", "int64", offset))
        return self

    def encode(self, param_order=None):
        body = _ld(1, self.name.encode()) + _ld(2, self.datatype.encode())
        if self.shape:
            body += _ld(3, b"".join(_varint(d) for d in self.shape))  # packed int64
        params = self.params if param_order is None else [p for k in param_order for p in self.params if p[0] == k]
        for key, kind, value in params:
            body += _map_entry(4, key, _infer_parameter(value, kind))
        return body


def grpc_request_bytes(model_name, inputs, model_version="", request_id="", outputs=None, parameters=None,
                       input_param_order=None):
    """grpc/_utils.py:80-139 -> serialized ModelInferRequest.

    Field order is the proto's field-number order (how protobuf serialises).  Map
    entry order is not defined by protobuf: ``parameters`` /
    ``input_param_order`` give the order to compare against; pinned golden cases
    use at most one entry per map.
    """
    out = b""
    if model_name:
        out += _ld(1, model_name.encode())
    if model_version:
        out += _ld(2, model_version.encode())
    if request_id:
        out += _ld(3, request_id.encode())
    for key, kind, value in (parameters or []):
        out += _map_entry(4, key, _infer_parameter(value, kind))
    for i in inputs:
        out += _ld(5, i.encode(input_param_order))
    for name in (outputs or []):
        out += _ld(6, _ld(1, name.encode()))
    for i in inputs:
        if i.raw is not None:
            out += _ld(7, i.raw)
    return out
