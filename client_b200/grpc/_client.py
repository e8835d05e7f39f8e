"""gRPC client.

Drop-in for ``tritonclient.grpc.InferenceServerClient`` (reference:
src/python/library/tritonclient/grpc/_client.py:57-1936): same constructor,
methods, arguments and error behaviour for health / metadata / repository /
statistics / trace / log / shared-memory control and for ``infer``,
``async_infer`` and the bidirectional stream (``start_stream`` /
``async_stream_infer`` / ``stop_stream``).
"""

import base64
import json
import os

import grpc
from google.protobuf.json_format import MessageToJson

from .._client import InferenceServerClientBase
from .._request import Request
from . import service_pb2, service_pb2_grpc
from ._infer_result import InferResult
from ._infer_stream import _InferStream, _RequestIterator
from ._utils import (
    _get_inference_request,
    _grpc_compression_type,
    get_cancelled_error,
    get_error_grpc,
    raise_error,
    raise_error_grpc,
)

INT32_MAX = 2**31 - 1
# gRPC caps a message at INT32_MAX bytes; larger tensors must go through shm
MAX_GRPC_MESSAGE_SIZE = INT32_MAX


class KeepAliveOptions:
    """HTTP/2 keepalive settings of the channel (reference :57-98).

    Parameters
    ----------
    keepalive_time_ms : int
        Period after which a keepalive ping is sent (default INT32_MAX: never).
    keepalive_timeout_ms : int
        How long the sender waits for the ping ack (default 20000).
    keepalive_permit_without_calls : bool
        Allow pings without any call in flight (default False).
    http2_max_pings_without_data : int
        Pings allowed without data frames (default 2).
    """

    def __init__(self, keepalive_time_ms=INT32_MAX, keepalive_timeout_ms=20000,
                 keepalive_permit_without_calls=False, http2_max_pings_without_data=2):
        self.keepalive_time_ms = keepalive_time_ms
        self.keepalive_timeout_ms = keepalive_timeout_ms
        self.keepalive_permit_without_calls = keepalive_permit_without_calls
        self.http2_max_pings_without_data = http2_max_pings_without_data


class CallContext:
    """Handle of an ``async_infer`` call; ``cancel()`` cancels it (reference :101-116)."""

    def __init__(self, grpc_future):
        self.__grpc_future = grpc_future

    def cancel(self):
        self.__grpc_future.cancel()


def _read_file(path):
    if path is None:
        return None
    with open(path, "rb") as fh:
        return fh.read()


class InferenceServerClient(InferenceServerClientBase):
    """Client of the inference server's gRPC endpoint.  Most methods are thread
    safe except ``infer`` / ``async_infer`` / the stream calls (one client per
    thread, reference :119-175).

    Parameters
    ----------
    url : str
        ``host:port`` of the server, e.g. ``localhost:8001``.
    verbose : bool
        Print requests and responses.
    ssl, root_certificates, private_key, certificate_chain, creds
        TLS settings as in the reference.
    keepalive_options : KeepAliveOptions
    channel_args : list of (key, value)
        Raw channel arguments; when given they replace the defaults entirely.
    """

    def __init__(self, url, verbose=False, ssl=False, root_certificates=None, private_key=None,
                 certificate_chain=None, creds=None, keepalive_options=None, channel_args=None, transport=None):
        super().__init__()
        # transport="native" (extension of this drop-in, default from TB200_GRPC_TRANSPORT): infer()
        # rides on libtb200client's own HTTP/2 channel instead of grpcio (cleartext only)
        transport = transport or os.environ.get("TB200_GRPC_TRANSPORT", "grpcio")
        if transport not in ("grpcio", "native"):
            raise_error("transport must be 'grpcio' or 'native'")
        self._native = None
        if transport == "native" and not (ssl or creds):
            from ._native_channel import NativeChannel

            self._native = NativeChannel(url)
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
            self._channel = grpc.secure_channel(url, creds, options=options)
        elif ssl:
            creds = grpc.ssl_channel_credentials(
                root_certificates=_read_file(root_certificates),
                private_key=_read_file(private_key),
                certificate_chain=_read_file(certificate_chain),
            )
            self._channel = grpc.secure_channel(url, creds, options=options)
        else:
            self._channel = grpc.insecure_channel(url, options=options)
        self._client_stub = service_pb2_grpc.GRPCInferenceServiceStub(self._channel)
        self._verbose = verbose
        self._stream = None

    def _get_metadata(self, headers):
        request = Request(headers)
        self._call_plugin(request)
        return request.headers.items() if request.headers is not None else ()

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        """Stop the stream, if any, and close the channel."""
        self.stop_stream()
        self._channel.close()
        if self._native is not None:
            self._native.close()
            self._native = None

    # one round trip of a unary control-plane rpc
    def _unary(self, rpc, request, headers, client_timeout, label=None):
        metadata = self._get_metadata(headers)
        if self._verbose:
            print("{}, metadata {}\n{}".format(label or rpc, metadata, request))
        try:
            response = getattr(self._client_stub, rpc)(request=request, metadata=metadata, timeout=client_timeout)
        except grpc.RpcError as rpc_error:
            raise_error_grpc(rpc_error)
        if self._verbose:
            print(response)
        return response

    @staticmethod
    def _maybe_json(response, as_json):
        if as_json:
            return json.loads(MessageToJson(response, preserving_proto_field_name=True))
        return response

    @staticmethod
    def _check_version(model_version):
        if type(model_version) != str:
            raise_error("model version must be a string")

    # -- health / metadata --------------------------------------------------------------
    def is_server_live(self, headers=None, client_timeout=None):
        """True when the server is live."""
        return self._unary("ServerLive", service_pb2.ServerLiveRequest(), headers, client_timeout, "is_server_live").live

    def is_server_ready(self, headers=None, client_timeout=None):
        """True when the server is ready."""
        return self._unary("ServerReady", service_pb2.ServerReadyRequest(), headers, client_timeout, "is_server_ready").ready

    def is_model_ready(self, model_name, model_version="", headers=None, client_timeout=None):
        """True when the model (version) is ready."""
        self._check_version(model_version)
        request = service_pb2.ModelReadyRequest(name=model_name, version=model_version)
        return self._unary("ModelReady", request, headers, client_timeout, "is_model_ready").ready

    def get_server_metadata(self, headers=None, as_json=False, client_timeout=None):
        """Server metadata (message, or dict with ``as_json``)."""
        r = self._unary("ServerMetadata", service_pb2.ServerMetadataRequest(), headers, client_timeout, "get_server_metadata")
        return self._maybe_json(r, as_json)

    def get_model_metadata(self, model_name, model_version="", headers=None, as_json=False, client_timeout=None):
        """Model metadata."""
        self._check_version(model_version)
        request = service_pb2.ModelMetadataRequest(name=model_name, version=model_version)
        return self._maybe_json(self._unary("ModelMetadata", request, headers, client_timeout, "get_model_metadata"), as_json)

    def get_model_config(self, model_name, model_version="", headers=None, as_json=False, client_timeout=None):
        """Model configuration."""
        self._check_version(model_version)
        request = service_pb2.ModelConfigRequest(name=model_name, version=model_version)
        return self._maybe_json(self._unary("ModelConfig", request, headers, client_timeout, "get_model_config"), as_json)

    def get_model_repository_index(self, headers=None, as_json=False, client_timeout=None):
        """Index of the model repository."""
        r = self._unary("RepositoryIndex", service_pb2.RepositoryIndexRequest(), headers, client_timeout, "get_model_repository_index")
        return self._maybe_json(r, as_json)

    def load_model(self, model_name, headers=None, config=None, files=None, client_timeout=None):
        """Ask the server to load (or reload) a model, optionally with a config
        override and override files."""
        request = service_pb2.RepositoryModelLoadRequest(model_name=model_name)
        if config is not None:
            request.parameters["config"].string_param = config
        for path, content in (files or {}).items():
            request.parameters[path].bytes_param = content
        self._unary("RepositoryModelLoad", request, headers, client_timeout, "load_model")
        if self._verbose:
            print("Loaded model '{}'".format(model_name))

    def unload_model(self, model_name, headers=None, unload_dependents=False, client_timeout=None):
        """Ask the server to unload a model."""
        request = service_pb2.RepositoryModelUnloadRequest(model_name=model_name)
        request.parameters["unload_dependents"].bool_param = unload_dependents
        self._unary("RepositoryModelUnload", request, headers, client_timeout, "unload_model")
        if self._verbose:
            print("Unloaded model '{}'".format(model_name))

    def get_inference_statistics(self, model_name="", model_version="", headers=None, as_json=False, client_timeout=None):
        """Inference statistics of one model (version) or of all models."""
        self._check_version(model_version)
        request = service_pb2.ModelStatisticsRequest(name=model_name, version=model_version)
        return self._maybe_json(self._unary("ModelStatistics", request, headers, client_timeout, "get_inference_statistics"), as_json)

    def update_trace_settings(self, model_name=None, settings={}, headers=None, as_json=False, client_timeout=None):
        """Update trace settings (a value of None clears a setting)."""
        request = service_pb2.TraceSettingRequest()
        if model_name is not None and model_name != "":
            request.model_name = model_name
        for key, value in settings.items():
            if value is None:
                request.settings[key]  # present but empty: clear
            else:
                request.settings[key].value.extend(value if isinstance(value, list) else [value])
        return self._maybe_json(self._unary("TraceSetting", request, headers, client_timeout, "update_trace_settings"), as_json)

    def get_trace_settings(self, model_name=None, headers=None, as_json=False, client_timeout=None):
        """Trace settings of a model, or the global ones."""
        request = service_pb2.TraceSettingRequest()
        if model_name is not None and model_name != "":
            request.model_name = model_name
        return self._maybe_json(self._unary("TraceSetting", request, headers, client_timeout, "get_trace_settings"), as_json)

    def update_log_settings(self, settings, headers=None, as_json=False, client_timeout=None):
        """Update the global log settings."""
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
        return self._maybe_json(self._unary("LogSettings", request, headers, client_timeout, "update_log_settings"), as_json)

    def get_log_settings(self, headers=None, as_json=False, client_timeout=None):
        """The global log settings."""
        r = self._unary("LogSettings", service_pb2.LogSettingsRequest(), headers, client_timeout, "get_log_settings")
        return self._maybe_json(r, as_json)

    # -- shared memory control plane --------------------------------------------------------
    def get_system_shared_memory_status(self, region_name="", headers=None, as_json=False, client_timeout=None):
        """Status of one / all registered system shared memory regions."""
        request = service_pb2.SystemSharedMemoryStatusRequest(name=region_name)
        return self._maybe_json(self._unary("SystemSharedMemoryStatus", request, headers, client_timeout, "get_system_shared_memory_status"), as_json)

    def register_system_shared_memory(self, name, key, byte_size, offset=0, headers=None, client_timeout=None):
        """Register a system shared memory region with the server."""
        request = service_pb2.SystemSharedMemoryRegisterRequest(name=name, key=key, offset=offset, byte_size=byte_size)
        self._unary("SystemSharedMemoryRegister", request, headers, client_timeout, "register_system_shared_memory")
        if self._verbose:
            print("Registered system shared memory with name '{}'".format(name))

    def unregister_system_shared_memory(self, name="", headers=None, client_timeout=None):
        """Unregister one region, or all when ``name`` is empty."""
        request = service_pb2.SystemSharedMemoryUnregisterRequest(name=name)
        self._unary("SystemSharedMemoryUnregister", request, headers, client_timeout, "unregister_system_shared_memory")
        if self._verbose:
            if name != "":
                print("Unregistered system shared memory with name '{}'".format(name))
            else:
                print("Unregistered all system shared memory regions")

    def get_cuda_shared_memory_status(self, region_name="", headers=None, as_json=False, client_timeout=None):
        """Status of one / all registered CUDA shared memory regions."""
        request = service_pb2.CudaSharedMemoryStatusRequest(name=region_name)
        return self._maybe_json(self._unary("CudaSharedMemoryStatus", request, headers, client_timeout, "get_cuda_shared_memory_status"), as_json)

    def register_cuda_shared_memory(self, name, raw_handle, device_id, byte_size, headers=None, client_timeout=None):
        """Register a CUDA shared memory region; ``raw_handle`` is the base64 IPC
        handle, sent as the raw 64 bytes (reference :1378-1393)."""
        request = service_pb2.CudaSharedMemoryRegisterRequest(
            name=name, raw_handle=base64.b64decode(raw_handle), device_id=device_id, byte_size=byte_size
        )
        self._unary("CudaSharedMemoryRegister", request, headers, client_timeout, "register_cuda_shared_memory")
        if self._verbose:
            print("Registered cuda shared memory with name '{}'".format(name))

    def unregister_cuda_shared_memory(self, name="", headers=None, client_timeout=None):
        """Unregister one region, or all when ``name`` is empty."""
        request = service_pb2.CudaSharedMemoryUnregisterRequest(name=name)
        self._unary("CudaSharedMemoryUnregister", request, headers, client_timeout, "unregister_cuda_shared_memory")
        if self._verbose:
            if name != "":
                print("Unregistered cuda shared memory with name '{}'".format(name))
            else:
                print("Unregistered all cuda shared memory regions")

    # -- inference ----------------------------------------------------------------------------
    def _build_request(self, model_name, inputs, model_version, outputs, request_id, sequence_id,
                       sequence_start, sequence_end, priority, timeout, parameters):
        self._check_version(model_version)
        return _get_inference_request(
            model_name=model_name, inputs=inputs, model_version=model_version, request_id=request_id,
            outputs=outputs, sequence_id=sequence_id, sequence_start=sequence_start,
            sequence_end=sequence_end, priority=priority, timeout=timeout, parameters=parameters,
        )

    def infer(self, model_name, inputs, model_version="", outputs=None, request_id="", sequence_id=0,
              sequence_start=False, sequence_end=False, priority=0, timeout=None, client_timeout=None,
              headers=None, compression_algorithm=None, parameters=None):
        """Run a synchronous inference; returns :py:class:`InferResult`."""
        metadata = self._get_metadata(headers)
        request = self._build_request(model_name, inputs, model_version, outputs, request_id, sequence_id,
                                      sequence_start, sequence_end, priority, timeout, parameters)
        if self._verbose:
            print("infer, metadata {}\n{}".format(metadata, request))
        if self._native is not None and compression_algorithm is None:
            raw = self._native.unary("/inference.GRPCInferenceService/ModelInfer", request.SerializeToString(), metadata, client_timeout)
            response = service_pb2.ModelInferResponse.FromString(raw)
            if self._verbose:
                print(response)
            return InferResult(response)
        try:
            response = self._client_stub.ModelInfer(
                request=request, metadata=metadata, timeout=client_timeout,
                compression=_grpc_compression_type(compression_algorithm),
            )
        except grpc.RpcError as rpc_error:
            raise_error_grpc(rpc_error)
        if self._verbose:
            print(response)
        return InferResult(response)

    def async_infer(self, model_name, inputs, callback, model_version="", outputs=None, request_id="",
                    sequence_id=0, sequence_start=False, sequence_end=False, priority=0, timeout=None,
                    client_timeout=None, headers=None, compression_algorithm=None, parameters=None):
        """Send an inference request without waiting; ``callback(result, error)`` runs
        when the response (or failure) arrives.  Returns a :py:class:`CallContext`."""

        def on_done(call_future):
            result = error = None
            try:
                response = call_future.result()
                if self._verbose:
                    print(response)
                result = InferResult(response)
            except grpc.RpcError as rpc_error:
                error = get_error_grpc(rpc_error)
            except grpc.FutureCancelledError:
                error = get_cancelled_error()
            callback(result=result, error=error)

        metadata = self._get_metadata(headers)
        request = self._build_request(model_name, inputs, model_version, outputs, request_id, sequence_id,
                                      sequence_start, sequence_end, priority, timeout, parameters)
        if self._verbose:
            print("async_infer, metadata {}\n{}".format(metadata, request))
        try:
            self._call_future = self._client_stub.ModelInfer.future(
                request=request, metadata=metadata, timeout=client_timeout,
                compression=_grpc_compression_type(compression_algorithm),
            )
            if self._verbose:
                message = "Sent request"
                if request_id != "":
                    message = message + " '{}'".format(request_id)
                print(message)
            self._call_future.add_done_callback(on_done)
            return CallContext(self._call_future)
        except grpc.RpcError as rpc_error:
            raise_error_grpc(rpc_error)

    def start_stream(self, callback, stream_timeout=None, headers=None, compression_algorithm=None):
        """Open the bidirectional inference stream; ``callback(result, error)`` runs on
        the reader thread for every response.  One stream per client."""
        if self._stream is not None:
            raise_error(
                "cannot start another stream with one already running. "
                "'InferenceServerClient' supports only a single active "
                "stream at a given time."
            )
        metadata = self._get_metadata(headers)
        self._stream = _InferStream(callback, self._verbose)
        try:
            response_iterator = self._client_stub.ModelStreamInfer(
                _RequestIterator(self._stream), metadata=metadata, timeout=stream_timeout,
                compression=_grpc_compression_type(compression_algorithm),
            )
            self._stream._init_handler(response_iterator)
        except grpc.RpcError as rpc_error:
            raise_error_grpc(rpc_error)

    def stop_stream(self, cancel_requests=False):
        """Close the stream (cancelling pending requests if asked)."""
        if self._stream is not None:
            self._stream.close(cancel_requests)
        self._stream = None

    def async_stream_infer(self, model_name, inputs, model_version="", outputs=None, request_id="",
                           sequence_id=0, sequence_start=False, sequence_end=False,
                           enable_empty_final_response=False, priority=0, timeout=None, parameters=None):
        """Queue an inference request on the active stream."""
        if self._stream is None:
            raise_error("stream not available, use start_stream() to make one available.")
        request = self._build_request(model_name, inputs, model_version, outputs, request_id, sequence_id,
                                      sequence_start, sequence_end, priority, timeout, parameters)
        if enable_empty_final_response:
            request.parameters["triton_enable_empty_final_response"].bool_param = True
        if self._verbose:
            print("async_stream_infer\n{}".format(request))
        self._stream._enqueue_request(request)
        if self._verbose:
            print("enqueued request {} to stream...".format(request_id))
