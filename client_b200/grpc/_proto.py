"""KServe-v2 / Triton gRPC schema built at import time (no protoc in this image).

The wire schema of ``inference.GRPCInferenceService`` is written here as a compact
table and turned into real protobuf message classes with
``descriptor_pb2.FileDescriptorProto`` + ``message_factory``.  Field names,
numbers and types follow the only copy of the schema in the reference,
src/rust/triton-client/proto/grpc_service.proto:40-218 (service), :226-1800
(messages), and model_config.proto for the whole of ModelConfig (generated table).  The
reference Python client gets the same classes from generated ``service_pb2`` /
``service_pb2_grpc`` / ``model_config_pb2`` modules
(src/python/library/build_wheel.py:110-139); the module objects built by
``build_modules()`` expose the same attribute names.
"""

import types

import grpc
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
_SCALARS = {
    "bool": _F.TYPE_BOOL, "int32": _F.TYPE_INT32, "int64": _F.TYPE_INT64,
    "uint32": _F.TYPE_UINT32, "uint64": _F.TYPE_UINT64, "float": _F.TYPE_FLOAT,
    "double": _F.TYPE_DOUBLE, "string": _F.TYPE_STRING, "bytes": _F.TYPE_BYTES,
}

# message -> list of fields: (name, number, type, label) with label in
# "" (singular) | "rep" (repeated) | "map:<valuetype>" | "oneof:<group>"
_PARAM_ONEOF = "oneof:parameter_choice"
SCHEMA = {
    "ServerLiveRequest": [],
    "ServerLiveResponse": [("live", 1, "bool", "")],
    "ServerReadyRequest": [],
    "ServerReadyResponse": [("ready", 1, "bool", "")],
    "ModelReadyRequest": [("name", 1, "string", ""), ("version", 2, "string", "")],
    "ModelReadyResponse": [("ready", 1, "bool", "")],
    "ServerMetadataRequest": [],
    "ServerMetadataResponse": [("name", 1, "string", ""), ("version", 2, "string", ""), ("extensions", 3, "string", "rep")],
    "ModelMetadataRequest": [("name", 1, "string", ""), ("version", 2, "string", "")],
    "ModelMetadataResponse": [
        ("name", 1, "string", ""), ("versions", 2, "string", "rep"), ("platform", 3, "string", ""),
        ("inputs", 4, ".TensorMetadata", "rep"), ("outputs", 5, ".TensorMetadata", "rep"),
    ],
    "ModelMetadataResponse.TensorMetadata": [("name", 1, "string", ""), ("datatype", 2, "string", ""), ("shape", 3, "int64", "rep")],
    "InferParameter": [
        ("bool_param", 1, "bool", _PARAM_ONEOF), ("int64_param", 2, "int64", _PARAM_ONEOF),
        ("string_param", 3, "string", _PARAM_ONEOF), ("double_param", 4, "double", _PARAM_ONEOF),
        ("uint64_param", 5, "uint64", _PARAM_ONEOF),
    ],
    "InferTensorContents": [
        ("bool_contents", 1, "bool", "rep"), ("int_contents", 2, "int32", "rep"), ("int64_contents", 3, "int64", "rep"),
        ("uint_contents", 4, "uint32", "rep"), ("uint64_contents", 5, "uint64", "rep"), ("fp32_contents", 6, "float", "rep"),
        ("fp64_contents", 7, "double", "rep"), ("bytes_contents", 8, "bytes", "rep"),
    ],
    "ModelInferRequest": [
        ("model_name", 1, "string", ""), ("model_version", 2, "string", ""), ("id", 3, "string", ""),
        ("parameters", 4, "InferParameter", "map"), ("inputs", 5, ".InferInputTensor", "rep"),
        ("outputs", 6, ".InferRequestedOutputTensor", "rep"), ("raw_input_contents", 7, "bytes", "rep"),
    ],
    "ModelInferRequest.InferInputTensor": [
        ("name", 1, "string", ""), ("datatype", 2, "string", ""), ("shape", 3, "int64", "rep"),
        ("parameters", 4, "InferParameter", "map"), ("contents", 5, "InferTensorContents", ""),
    ],
    "ModelInferRequest.InferRequestedOutputTensor": [("name", 1, "string", ""), ("parameters", 2, "InferParameter", "map")],
    "ModelInferResponse": [
        ("model_name", 1, "string", ""), ("model_version", 2, "string", ""), ("id", 3, "string", ""),
        ("parameters", 4, "InferParameter", "map"), ("outputs", 5, ".InferOutputTensor", "rep"),
        ("raw_output_contents", 6, "bytes", "rep"),
    ],
    "ModelInferResponse.InferOutputTensor": [
        ("name", 1, "string", ""), ("datatype", 2, "string", ""), ("shape", 3, "int64", "rep"),
        ("parameters", 4, "InferParameter", "map"), ("contents", 5, "InferTensorContents", ""),
    ],
    "ModelStreamInferResponse": [("error_message", 1, "string", ""), ("infer_response", 2, "ModelInferResponse", "")],
    "ModelConfigRequest": [("name", 1, "string", ""), ("version", 2, "string", "")],
    "ModelConfigResponse": [("config", 1, "ModelConfig", "")],
    "ModelStatisticsRequest": [("name", 1, "string", ""), ("version", 2, "string", "")],
    "StatisticDuration": [("count", 1, "uint64", ""), ("ns", 2, "uint64", "")],
    "InferStatistics": [
        ("success", 1, "StatisticDuration", ""), ("fail", 2, "StatisticDuration", ""), ("queue", 3, "StatisticDuration", ""),
        ("compute_input", 4, "StatisticDuration", ""), ("compute_infer", 5, "StatisticDuration", ""),
        ("compute_output", 6, "StatisticDuration", ""), ("cache_hit", 7, "StatisticDuration", ""),
        ("cache_miss", 8, "StatisticDuration", ""),
    ],
    "InferResponseStatistics": [
        ("compute_infer", 1, "StatisticDuration", ""), ("compute_output", 2, "StatisticDuration", ""),
        ("success", 3, "StatisticDuration", ""), ("fail", 4, "StatisticDuration", ""),
        ("empty_response", 5, "StatisticDuration", ""), ("cancel", 6, "StatisticDuration", ""),
    ],
    "InferBatchStatistics": [
        ("batch_size", 1, "uint64", ""), ("compute_input", 2, "StatisticDuration", ""),
        ("compute_infer", 3, "StatisticDuration", ""), ("compute_output", 4, "StatisticDuration", ""),
    ],
    "MemoryUsage": [("type", 1, "string", ""), ("id", 2, "int64", ""), ("byte_size", 3, "uint64", "")],
    "ModelStatistics": [
        ("name", 1, "string", ""), ("version", 2, "string", ""), ("last_inference", 3, "uint64", ""),
        ("inference_count", 4, "uint64", ""), ("execution_count", 5, "uint64", ""),
        ("inference_stats", 6, "InferStatistics", ""), ("batch_stats", 7, "InferBatchStatistics", "rep"),
        ("memory_usage", 8, "MemoryUsage", "rep"), ("response_stats", 9, "InferResponseStatistics", "map"),
    ],
    "ModelStatisticsResponse": [("model_stats", 1, "ModelStatistics", "rep")],
    "ModelRepositoryParameter": [
        ("bool_param", 1, "bool", _PARAM_ONEOF), ("int64_param", 2, "int64", _PARAM_ONEOF),
        ("string_param", 3, "string", _PARAM_ONEOF), ("bytes_param", 4, "bytes", _PARAM_ONEOF),
    ],
    "RepositoryIndexRequest": [("repository_name", 1, "string", ""), ("ready", 2, "bool", "")],
    "RepositoryIndexResponse": [("models", 1, ".ModelIndex", "rep")],
    "RepositoryIndexResponse.ModelIndex": [("name", 1, "string", ""), ("version", 2, "string", ""), ("state", 3, "string", ""), ("reason", 4, "string", "")],
    "RepositoryModelLoadRequest": [("repository_name", 1, "string", ""), ("model_name", 2, "string", ""), ("parameters", 3, "ModelRepositoryParameter", "map")],
    "RepositoryModelLoadResponse": [],
    "RepositoryModelUnloadRequest": [("repository_name", 1, "string", ""), ("model_name", 2, "string", ""), ("parameters", 3, "ModelRepositoryParameter", "map")],
    "RepositoryModelUnloadResponse": [],
    "SystemSharedMemoryStatusRequest": [("name", 1, "string", "")],
    "SystemSharedMemoryStatusResponse": [("regions", 1, ".RegionStatus", "map")],
    "SystemSharedMemoryStatusResponse.RegionStatus": [("name", 1, "string", ""), ("key", 2, "string", ""), ("offset", 3, "uint64", ""), ("byte_size", 4, "uint64", "")],
    "SystemSharedMemoryRegisterRequest": [("name", 1, "string", ""), ("key", 2, "string", ""), ("offset", 3, "uint64", ""), ("byte_size", 4, "uint64", "")],
    "SystemSharedMemoryRegisterResponse": [],
    "SystemSharedMemoryUnregisterRequest": [("name", 1, "string", "")],
    "SystemSharedMemoryUnregisterResponse": [],
    "CudaSharedMemoryStatusRequest": [("name", 1, "string", "")],
    "CudaSharedMemoryStatusResponse": [("regions", 1, ".RegionStatus", "map")],
    "CudaSharedMemoryStatusResponse.RegionStatus": [("name", 1, "string", ""), ("device_id", 2, "uint64", ""), ("byte_size", 3, "uint64", "")],
    "CudaSharedMemoryRegisterRequest": [("name", 1, "string", ""), ("raw_handle", 2, "bytes", ""), ("device_id", 3, "int64", ""), ("byte_size", 4, "uint64", "")],
    "CudaSharedMemoryRegisterResponse": [],
    "CudaSharedMemoryUnregisterRequest": [("name", 1, "string", "")],
    "CudaSharedMemoryUnregisterResponse": [],
    "TraceSettingRequest": [("settings", 1, ".SettingValue", "map"), ("model_name", 2, "string", "")],
    "TraceSettingRequest.SettingValue": [("value", 1, "string", "rep")],
    "TraceSettingResponse": [("settings", 1, ".SettingValue", "map")],
    "TraceSettingResponse.SettingValue": [("value", 1, "string", "rep")],
    "LogSettingsRequest": [("settings", 1, ".SettingValue", "map")],
    "LogSettingsRequest.SettingValue": [("bool_param", 1, "bool", _PARAM_ONEOF), ("uint32_param", 2, "uint32", _PARAM_ONEOF), ("string_param", 3, "string", _PARAM_ONEOF)],
    "LogSettingsResponse": [("settings", 1, ".SettingValue", "map")],
    "LogSettingsResponse.SettingValue": [("bool_param", 1, "bool", _PARAM_ONEOF), ("uint32_param", 2, "uint32", _PARAM_ONEOF), ("string_param", 3, "string", _PARAM_ONEOF)],
}

# model_config.proto: the full ModelConfig schema (55 messages, 9 enums) lives in the generated
# table _model_config_schema.py (oracle/gen_proto_fixture.py --emit-model-config); qualified type
# names, 'map:<key type>' labels.
from ._model_config_schema import MODEL_CONFIG_ENUMS, MODEL_CONFIG_SCHEMA  # noqa: E402

DATA_TYPE_ENUM = MODEL_CONFIG_ENUMS["DataType"]
FORMAT_ENUM = MODEL_CONFIG_ENUMS["ModelInput.Format"]

# rpc name -> (request, response, client_streaming, server_streaming)
SERVICE = {
    "ServerLive": ("ServerLiveRequest", "ServerLiveResponse", False, False),
    "ServerReady": ("ServerReadyRequest", "ServerReadyResponse", False, False),
    "ModelReady": ("ModelReadyRequest", "ModelReadyResponse", False, False),
    "ServerMetadata": ("ServerMetadataRequest", "ServerMetadataResponse", False, False),
    "ModelMetadata": ("ModelMetadataRequest", "ModelMetadataResponse", False, False),
    "ModelInfer": ("ModelInferRequest", "ModelInferResponse", False, False),
    "ModelStreamInfer": ("ModelInferRequest", "ModelStreamInferResponse", True, True),
    "ModelConfig": ("ModelConfigRequest", "ModelConfigResponse", False, False),
    "ModelStatistics": ("ModelStatisticsRequest", "ModelStatisticsResponse", False, False),
    "RepositoryIndex": ("RepositoryIndexRequest", "RepositoryIndexResponse", False, False),
    "RepositoryModelLoad": ("RepositoryModelLoadRequest", "RepositoryModelLoadResponse", False, False),
    "RepositoryModelUnload": ("RepositoryModelUnloadRequest", "RepositoryModelUnloadResponse", False, False),
    "SystemSharedMemoryStatus": ("SystemSharedMemoryStatusRequest", "SystemSharedMemoryStatusResponse", False, False),
    "SystemSharedMemoryRegister": ("SystemSharedMemoryRegisterRequest", "SystemSharedMemoryRegisterResponse", False, False),
    "SystemSharedMemoryUnregister": ("SystemSharedMemoryUnregisterRequest", "SystemSharedMemoryUnregisterResponse", False, False),
    "CudaSharedMemoryStatus": ("CudaSharedMemoryStatusRequest", "CudaSharedMemoryStatusResponse", False, False),
    "CudaSharedMemoryRegister": ("CudaSharedMemoryRegisterRequest", "CudaSharedMemoryRegisterResponse", False, False),
    "CudaSharedMemoryUnregister": ("CudaSharedMemoryUnregisterRequest", "CudaSharedMemoryUnregisterResponse", False, False),
    "TraceSetting": ("TraceSettingRequest", "TraceSettingResponse", False, False),
    "LogSettings": ("LogSettingsRequest", "LogSettingsResponse", False, False),
}
SERVICE_NAME = "inference.GRPCInferenceService"
PACKAGE = "inference"


def _camel(name):
    return "".join(p.capitalize() for p in name.split("_"))


def _resolve(type_name, owner):
    """'.Nested' is relative to the owning top-level message; anything else is a qualified name."""
    if type_name.startswith("."):
        return ".%s.%s%s" % (PACKAGE, owner.split(".")[0], type_name)
    return ".%s.%s" % (PACKAGE, type_name)


def _json_name(fname):
    parts = fname.split("_")
    return parts[0] + "".join(p.capitalize() for p in parts[1:])


def _fill_message(msg_proto, full_name, fields, schema, enums):
    oneofs = {}
    for fname, number, ftype, label in fields:
        f = msg_proto.field.add()
        f.name, f.number = fname, number
        f.json_name = _json_name(fname)
        f.label = _F.LABEL_OPTIONAL
        if label == "map" or label.startswith("map:"):
            # map<K, V> = repeated nested <Name>Entry {K key=1; V value=2}
            entry = msg_proto.nested_type.add()
            entry.name = _camel(fname) + "Entry"
            entry.options.map_entry = True
            k = entry.field.add()
            k.name, k.number, k.label, k.json_name = "key", 1, _F.LABEL_OPTIONAL, "key"
            k.type = _SCALARS[label.split(":", 1)[1] if ":" in label else "string"]
            v = entry.field.add()
            v.name, v.number, v.label, v.json_name = "value", 2, _F.LABEL_OPTIONAL, "value"
            _set_type(v, ftype, full_name)
            f.label = _F.LABEL_REPEATED
            f.type = _F.TYPE_MESSAGE
            f.type_name = ".%s.%s.%s" % (PACKAGE, full_name, entry.name)
            continue
        if label == "rep":
            f.label = _F.LABEL_REPEATED
        elif label.startswith("oneof:"):
            group = label.split(":", 1)[1]
            if group not in oneofs:
                oneofs[group] = len(msg_proto.oneof_decl)
                msg_proto.oneof_decl.add().name = group
            f.oneof_index = oneofs[group]
        _set_type(f, ftype, full_name)
    prefix = full_name + "."
    for child, child_fields in schema.items():
        if child.startswith(prefix) and "." not in child[len(prefix):]:
            nested = msg_proto.nested_type.add()
            nested.name = child[len(prefix):]
            _fill_message(nested, child, child_fields, schema, enums)
    for ename, values in enums.items():
        if ename.startswith(prefix) and "." not in ename[len(prefix):]:
            _fill_enum(msg_proto.enum_type.add(), ename[len(prefix):], values)


def _fill_enum(enum_proto, name, values):
    enum_proto.name = name
    for n, v in values:
        ev = enum_proto.value.add()
        ev.name, ev.number = n, v


def _set_type(field, type_name, owner):
    if type_name in _SCALARS:
        field.type = _SCALARS[type_name]
    elif type_name.startswith("enum:"):
        field.type = _F.TYPE_ENUM
        field.type_name = _resolve(type_name[5:], owner)
    else:
        field.type = _F.TYPE_MESSAGE
        field.type_name = _resolve(type_name, owner)


def _file_descriptor():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "client_b200/grpc_service.proto"
    fd.package = PACKAGE
    fd.syntax = "proto3"
    for ename, values in MODEL_CONFIG_ENUMS.items():
        if "." not in ename:
            _fill_enum(fd.enum_type.add(), ename, values)
    for schema, enums in ((MODEL_CONFIG_SCHEMA, MODEL_CONFIG_ENUMS), (SCHEMA, {})):
        for name, fields in schema.items():
            if "." in name:
                continue
            m = fd.message_type.add()
            m.name = name
            _fill_message(m, name, fields, schema, enums)
    svc = fd.service.add()
    svc.name = "GRPCInferenceService"
    for rpc, (req, resp, cs, ss) in SERVICE.items():
        meth = svc.method.add()
        meth.name = rpc
        meth.input_type = ".%s.%s" % (PACKAGE, req)
        meth.output_type = ".%s.%s" % (PACKAGE, resp)
        meth.client_streaming, meth.server_streaming = cs, ss
    return fd


_modules = None


def _enum_wrapper(desc):
    try:
        from google.protobuf.internal import enum_type_wrapper

        return enum_type_wrapper.EnumTypeWrapper(desc)
    except Exception:  # pragma: no cover
        return desc


def build_modules():
    """(service_pb2, service_pb2_grpc, model_config_pb2) module objects."""
    global _modules
    if _modules is not None:
        return _modules
    pool = descriptor_pool.DescriptorPool()
    file_desc = pool.Add(_file_descriptor()) if hasattr(pool, "Add") else None
    if file_desc is None:
        file_desc = pool.FindFileByName("client_b200/grpc_service.proto")
    else:
        file_desc = pool.FindFileByName("client_b200/grpc_service.proto")

    def cls(name):
        return message_factory.GetMessageClass(pool.FindMessageTypeByName("%s.%s" % (PACKAGE, name)))

    service_pb2 = types.ModuleType("client_b200.grpc.service_pb2")
    service_pb2.DESCRIPTOR = file_desc
    for name in SCHEMA:
        if "." not in name:
            setattr(service_pb2, name, cls(name))
    model_config_pb2 = types.ModuleType("client_b200.grpc.model_config_pb2")
    model_config_pb2.DESCRIPTOR = file_desc
    for name in MODEL_CONFIG_SCHEMA:
        if "." not in name:
            setattr(model_config_pb2, name, cls(name))
    for ename, values in MODEL_CONFIG_ENUMS.items():
        if "." not in ename:  # top-level enum: its values are module attributes, as in generated code
            setattr(model_config_pb2, ename, _enum_wrapper(pool.FindEnumTypeByName("%s.%s" % (PACKAGE, ename))))
            for n, v in values:
                setattr(model_config_pb2, n, v)
    service_pb2.ModelConfig = model_config_pb2.ModelConfig

    service_pb2_grpc = types.ModuleType("client_b200.grpc.service_pb2_grpc")

    class GRPCInferenceServiceStub:
        """Client stub: one multicallable per rpc of grpc_service.proto:40-218."""

        def __init__(self, channel):
            for rpc, (req, resp, cs, ss) in SERVICE.items():
                factory = channel.stream_stream if (cs and ss) else channel.unary_unary
                setattr(
                    self,
                    rpc,
                    factory(
                        "/%s/%s" % (SERVICE_NAME, rpc),
                        request_serializer=getattr(service_pb2, req).SerializeToString,
                        response_deserializer=getattr(service_pb2, resp).FromString,
                    ),
                )

    class GRPCInferenceServiceServicer:
        """Server base class: every rpc answers UNIMPLEMENTED until overridden."""

    def _unimplemented(rpc):
        def handler(self, request, context):
            context.set_code(grpc.StatusCode.UNIMPLEMENTED)
            context.set_details("Method %s not implemented!" % rpc)
            raise NotImplementedError("Method %s not implemented!" % rpc)

        return handler

    for rpc in SERVICE:
        setattr(GRPCInferenceServiceServicer, rpc, _unimplemented(rpc))

    def add_GRPCInferenceServiceServicer_to_server(servicer, server):
        handlers = {}
        for rpc, (req, resp, cs, ss) in SERVICE.items():
            make = grpc.stream_stream_rpc_method_handler if (cs and ss) else grpc.unary_unary_rpc_method_handler
            handlers[rpc] = make(
                getattr(servicer, rpc),
                request_deserializer=getattr(service_pb2, req).FromString,
                response_serializer=getattr(service_pb2, resp).SerializeToString,
            )
        server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE_NAME, handlers),))

    service_pb2_grpc.GRPCInferenceServiceStub = GRPCInferenceServiceStub
    service_pb2_grpc.GRPCInferenceServiceServicer = GRPCInferenceServiceServicer
    service_pb2_grpc.add_GRPCInferenceServiceServicer_to_server = add_GRPCInferenceServiceServicer_to_server
    _modules = (service_pb2, service_pb2_grpc, model_config_pb2)
    return _modules
