"""asyncio gRPC client.

Drop-in for ``tritonclient.grpc.aio.InferenceServerClient`` (reference:
src/python/library/tritonclient/grpc/aio/__init__.py:51-810): coroutine twins of the
control-plane calls and ``infer``, and ``stream_infer`` (async iterator of request
dicts in, async iterator of ``(result, error)`` out).  Request assembly is shared with
the synchronous client.
"""

import base64
import json

import grpc
from google.protobuf.json_format import MessageToJson

from ..._client import InferenceServerClientBase
from ..._request import Request
from ...utils import InferenceServerException
from .. import InferInput, InferRequestedOutput, KeepAliveOptions, service_pb2, service_pb2_grpc  # noqa: F401
from .._client import MAX_GRPC_MESSAGE_SIZE, _read_file
from .._infer_result import InferResult
from .._utils import _get_inference_request, _grpc_compression_type, raise_error, raise_error_grpc

_STREAM_DEFAULTS = {
    "model_version": "", "outputs": None, "request_id": "", "sequence_id": 0, "sequence_start": False,
    "sequence_end": False, "priority": 0, "timeout": None, "parameters": None,
}


class InferenceServerClient(InferenceServerClientBase):
    """asyncio twin of :py:class:`client_b200.grpc.InferenceServerClient`; most calls
    are safe to issue concurrently from one event loop.  Constructor arguments as in
    the synchronous client."""

    def __init__(self, url, verbose=False, ssl=False, root_certificates=None, private_key=None,
                 certificate_chain=None, creds=None, keepalive_options=None, channel_args=None):
        super().__init__()
        if channel_args is not None:
            options = channel_args
        else:
            ka = keepalive_options or KeepAliveOptions()
            options = [
                ("grpc.max_send_message_length", MAX_GRPC_MESSAGE_SIZE),
                ("grpc.max_receive_message_length", MAX_GRPC_MESSAGE_SIZE),
                ("grpc.keepalive_time_ms", ka.keepalive_time_ms),
                ("grpc.keepalive_timeout_ms", ka.keepalive_timeout_ms),
                ("grpc.keepalive_permit_without_calls", ka.keepalive_permit_without_calls),
                ("grpc.http2.max_pings_without_data", ka.http2_max_pings_without_data),
            ]
        if creds:
            self._channel = grpc.aio.secure_channel(url, creds, options=options)
        elif ssl:
            creds = grpc.ssl_channel_credentials(
                root_certificates=_read_file(root_certificates), private_key=_read_file(private_key),
                certificate_chain=_read_file(certificate_chain),
            )
            self._channel = grpc.aio.secure_channel(url, creds, options=options)
        else:
            self._channel = grpc.aio.insecure_channel(url, options=options)
        self._client_stub = service_pb2_grpc.GRPCInferenceServiceStub(self._channel)
        self._verbose = verbose

    def _return_response(self, response, as_json):
        if as_json:
            return json.loads(MessageToJson(response, preserving_proto_field_name=True))
        return response

    async def __aenter__(self):
        return self

    async def __aexit__(self, type, value, traceback):
        await self.close()

    async def close(self):
        """Close the channel."""
        await self._channel.close()

    def _get_metadata(self, headers):
        request = Request(headers)
        self._call_plugin(request)
        return request.headers.items() if request.headers is not None else ()

    async def _unary(self, rpc, request, headers, client_timeout):
        metadata = self._get_metadata(headers)
        if self._verbose:
            print("{}, metadata {}\n{}".format(rpc, metadata, request))
        try:
            response = await getattr(self._client_stub, rpc)(request=request, metadata=metadata, timeout=client_timeout)
        except grpc.RpcError as rpc_error:
            raise_error_grpc(rpc_error)
        if self._verbose:
            print(response)
        return response

    @staticmethod
    def _check_version(model_version):
        if type(model_version) != str:
            raise_error("model version must be a string")

    async def is_server_live(self, headers=None, client_timeout=None):
        return (await self._unary("ServerLive", service_pb2.ServerLiveRequest(), headers, client_timeout)).live

    async def is_server_ready(self, headers=None, client_timeout=None):
        return (await self._unary("ServerReady", service_pb2.ServerReadyRequest(), headers, client_timeout)).ready

    async def is_model_ready(self, model_name, model_version="", headers=None, client_timeout=None):
        self._check_version(model_version)
        request = service_pb2.ModelReadyRequest(name=model_name, version=model_version)
        return (await self._unary("ModelReady", request, headers, client_timeout)).ready

    async def get_server_metadata(self, headers=None, as_json=False, client_timeout=None):
        return self._return_response(await self._unary("ServerMetadata", service_pb2.ServerMetadataRequest(), headers, client_timeout), as_json)

    async def get_model_metadata(self, model_name, model_version="", headers=None, as_json=False, client_timeout=None):
        self._check_version(model_version)
        request = service_pb2.ModelMetadataRequest(name=model_name, version=model_version)
        return self._return_response(await self._unary("ModelMetadata", request, headers, client_timeout), as_json)

    async def get_model_config(self, model_name, model_version="", headers=None, as_json=False, client_timeout=None):
        self._check_version(model_version)
        request = service_pb2.ModelConfigRequest(name=model_name, version=model_version)
        return self._return_response(await self._unary("ModelConfig", request, headers, client_timeout), as_json)

    async def get_model_repository_index(self, headers=None, as_json=False, client_timeout=None):
        return self._return_response(await self._unary("RepositoryIndex", service_pb2.RepositoryIndexRequest(), headers, client_timeout), as_json)

    async def load_model(self, model_name, headers=None, config=None, files=None, client_timeout=None):
        request = service_pb2.RepositoryModelLoadRequest(model_name=model_name)
        if config is not None:
            request.parameters["config"].string_param = config
        for path, content in (files or {}).items():
            request.parameters[path].bytes_param = content
        await self._unary("RepositoryModelLoad", request, headers, client_timeout)

    async def unload_model(self, model_name, headers=None, unload_dependents=False, client_timeout=None):
        request = service_pb2.RepositoryModelUnloadRequest(model_name=model_name)
        request.parameters["unload_dependents"].bool_param = unload_dependents
        await self._unary("RepositoryModelUnload", request, headers, client_timeout)

    async def get_inference_statistics(self, model_name="", model_version="", headers=None, as_json=False, client_timeout=None):
        self._check_version(model_version)
        request = service_pb2.ModelStatisticsRequest(name=model_name, version=model_version)
        return self._return_response(await self._unary("ModelStatistics", request, headers, client_timeout), as_json)

    async def update_trace_settings(self, model_name=None, settings={}, headers=None, as_json=False, client_timeout=None):
        request = service_pb2.TraceSettingRequest()
        if model_name is not None and model_name != "":
            request.model_name = model_name
        for key, value in settings.items():
            if value is None:
                request.settings[key]
            else:
                request.settings[key].value.extend(value if isinstance(value, list) else [value])
        return self._return_response(await self._unary("TraceSetting", request, headers, client_timeout), as_json)

    async def get_trace_settings(self, model_name=None, headers=None, as_json=False, client_timeout=None):
        request = service_pb2.TraceSettingRequest()
        if model_name is not None and model_name != "":
            request.model_name = model_name
        return self._return_response(await self._unary("TraceSetting", request, headers, client_timeout), as_json)

    async def update_log_settings(self, settings, headers=None, as_json=False, client_timeout=None):
        request = service_pb2.LogSettingsRequest()
        for key, value in settings.items():
            if value is None:
                request.settings[key]
            elif key == "log_file" or key == "log_format":
                request.settings[key].string_param = value
            elif key == "log_verbose_level":
                request.settings[key].uint32_param = value
            else:
                request.settings[key].bool_param = value
        return self._return_response(await self._unary("LogSettings", request, headers, client_timeout), as_json)

    async def get_log_settings(self, headers=None, as_json=False, client_timeout=None):
        return self._return_response(await self._unary("LogSettings", service_pb2.LogSettingsRequest(), headers, client_timeout), as_json)

    async def get_system_shared_memory_status(self, region_name="", headers=None, as_json=False, client_timeout=None):
        request = service_pb2.SystemSharedMemoryStatusRequest(name=region_name)
        return self._return_response(await self._unary("SystemSharedMemoryStatus", request, headers, client_timeout), as_json)

    async def register_system_shared_memory(self, name, key, byte_size, offset=0, headers=None, client_timeout=None):
        request = service_pb2.SystemSharedMemoryRegisterRequest(name=name, key=key, offset=offset, byte_size=byte_size)
        await self._unary("SystemSharedMemoryRegister", request, headers, client_timeout)

    async def unregister_system_shared_memory(self, name="", headers=None, client_timeout=None):
        await self._unary("SystemSharedMemoryUnregister", service_pb2.SystemSharedMemoryUnregisterRequest(name=name), headers, client_timeout)

    async def get_cuda_shared_memory_status(self, region_name="", headers=None, as_json=False, client_timeout=None):
        request = service_pb2.CudaSharedMemoryStatusRequest(name=region_name)
        return self._return_response(await self._unary("CudaSharedMemoryStatus", request, headers, client_timeout), as_json)

    async def register_cuda_shared_memory(self, name, raw_handle, device_id, byte_size, headers=None, client_timeout=None):
        request = service_pb2.CudaSharedMemoryRegisterRequest(
            name=name, raw_handle=base64.b64decode(raw_handle), device_id=device_id, byte_size=byte_size
        )
        await self._unary("CudaSharedMemoryRegister", request, headers, client_timeout)

    async def unregister_cuda_shared_memory(self, name="", headers=None, client_timeout=None):
        await self._unary("CudaSharedMemoryUnregister", service_pb2.CudaSharedMemoryUnregisterRequest(name=name), headers, client_timeout)

    async def infer(self, model_name, inputs, model_version="", outputs=None, request_id="", sequence_id=0,
                    sequence_start=False, sequence_end=False, priority=0, timeout=None, client_timeout=None,
                    headers=None, compression_algorithm=None, parameters=None):
        """Run an inference; returns :py:class:`client_b200.grpc.InferResult`."""
        metadata = self._get_metadata(headers)
        self._check_version(model_version)
        request = _get_inference_request(
            model_name=model_name, inputs=inputs, model_version=model_version, request_id=request_id,
            outputs=outputs, sequence_id=sequence_id, sequence_start=sequence_start, sequence_end=sequence_end,
            priority=priority, timeout=timeout, parameters=parameters,
        )
        if self._verbose:
            print("infer, metadata {}\n{}".format(metadata, request))
        try:
            response = await self._client_stub.ModelInfer(
                request=request, metadata=metadata, timeout=client_timeout,
                compression=_grpc_compression_type(compression_algorithm),
            )
        except grpc.RpcError as rpc_error:
            raise_error_grpc(rpc_error)
        if self._verbose:
            print(response)
        return InferResult(response)

    def stream_infer(self, inputs_iterator, stream_timeout=None, headers=None, compression_algorithm=None):
        """Bidirectional streaming inference.

        ``inputs_iterator`` is an async iterator of dicts holding the arguments of
        ``async_stream_infer`` (``model_name`` and ``inputs`` required).  Returns an async
        iterator of ``(InferResult, InferenceServerException)`` tuples with a
        ``cancel()`` method (reference :688-810).
        """
        metadata = self._get_metadata(headers)

        async def requests():
            async for item in inputs_iterator:
                if type(item) != dict:
                    raise_error("inputs_iterator is not yielding a dict")
                if "model_name" not in item or "inputs" not in item:
                    raise_error("model_name and/or inputs is missing from inputs_iterator's yielded dict")
                args = dict(_STREAM_DEFAULTS)
                args.update(item)
                if type(args["model_version"]) != str:
                    raise_error("model_version must be a string")
                flag = args.pop("enable_empty_final_response", False)
                request = _get_inference_request(**args)
                if flag:
                    request.parameters["triton_enable_empty_final_response"].bool_param = True
                yield request

        verbose = self._verbose

        class _ResponseIterator:
            def __init__(self, call):
                self._call = call
                self._it = call.__aiter__()

            def __aiter__(self):
                return self

            async def __anext__(self):
                response = await self._it.__anext__()
                if verbose:
                    print(response)
                if response.error_message != "":
                    return None, InferenceServerException(msg=response.error_message)
                return InferResult(response.infer_response), None

            def cancel(self):
                return self._call.cancel()

        try:
            call = self._client_stub.ModelStreamInfer(
                requests(), metadata=metadata, timeout=stream_timeout,
                compression=_grpc_compression_type(compression_algorithm),
            )
            return _ResponseIterator(call)
        except grpc.RpcError as rpc_error:
            raise_error_grpc(rpc_error)
