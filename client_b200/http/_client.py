"""HTTP/REST (KServe-v2) client.

Drop-in for ``tritonclient.http.InferenceServerClient`` / ``InferAsyncRequest``
(reference: src/python/library/tritonclient/http/_client.py:44-1659): same
constructor, methods, arguments, URIs, request bodies and error behaviour.

Transport: the reference rides on gevent + geventhttpclient (neither is in this
image); here a pool of persistent raw HTTP/1.1 socket connections and a
thread pool carry the requests.  ``async_infer`` therefore has no
``gevent.sleep(0.01)`` floor (reference :1648-1651).
"""

import base64
import gzip
import json
import queue
import socket
import ssl as _ssl
import zlib
from concurrent.futures import ThreadPoolExecutor
from concurrent.futures import TimeoutError as _FutureTimeout
from urllib.parse import quote

from .._client import InferenceServerClientBase
from .._request import Request
from ..utils import raise_error
from ._infer_result import InferResult, _BufferResponse
from ._utils import _dumps, _get_inference_request, _get_query_string, _raise_if_error


class _HttpResponse(_BufferResponse):
    """Fully-read HTTP response with the accessors the result classes use
    (``status_code``, ``get(header)``, ``read(length)``)."""

    def __init__(self, status_code, headers, body):
        super().__init__(body, None)
        self.status_code = status_code
        self._lower = {k.lower(): v for k, v in headers}

    def get(self, key):
        return self._lower.get(key.lower())

    def __str__(self):
        return "<HTTP %d, %d bytes>" % (self.status_code, len(self._body))


class _ConnectionDropped(Exception):
    """The peer closed a keep-alive connection before answering."""


class _RawConnection:
    """One persistent HTTP/1.1 connection over a plain (or TLS) socket.

    ``http.client`` spends ~60 us per exchange in header bookkeeping
    (email.feedparser); requests here are one ``sendmsg`` of pre-joined header bytes plus
    the body, responses are split with ``bytes.find``."""

    def __init__(self, host, port, connect_timeout, network_timeout, ssl_context):
        sock = socket.create_connection((host, port), timeout=connect_timeout)
        sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        if ssl_context is not None:
            sock = ssl_context.wrap_socket(sock, server_hostname=host)
        sock.settimeout(network_timeout)
        self._sock = sock
        self._plain = ssl_context is None
        self._buf = b""
        self._host_only = ("Host: %s:%d\r\n" % (host, port)).encode("ascii")
        self._host_line = self._host_only + b"Accept-Encoding: identity\r\n"

    def close(self):
        try:
            self._sock.close()
        except OSError:
            pass

    def _more(self):
        chunk = self._sock.recv(1 << 18)
        if not chunk:
            raise _ConnectionDropped()
        self._buf += chunk

    def exchange(self, method, uri, body, headers):
        # the default "identity" only when the caller asks for no encoding of its own (one Accept-Encoding per request)
        own = any(k.lower() == "accept-encoding" for k in headers)
        head = [("%s %s HTTP/1.1\r\n" % (method, uri)).encode("ascii"), self._host_only if own else self._host_line]
        for k, v in headers.items():
            head.append(("%s: %s\r\n" % (k, v)).encode("latin-1"))
        body = body or b""
        if body or method == "POST":
            head.append(b"Content-Length: %d\r\n" % len(body))
        head.append(b"\r\n")
        head = b"".join(head)
        try:
            if len(body) <= 65536:
                self._sock.sendall(head + body)
            elif self._plain:
                sent = self._sock.sendmsg([head, body])
                if sent < len(head) + len(body):
                    rest = (head + bytes(body))[sent:] if sent < len(head) else memoryview(body)[sent - len(head):]
                    self._sock.sendall(rest)
            else:
                self._sock.sendall(head)
                self._sock.sendall(body)
        except (BrokenPipeError, ConnectionResetError):
            raise _ConnectionDropped()
        # ---- response
        try:
            while True:
                end = self._buf.find(b"\r\n\r\n")
                if end >= 0:
                    break
                self._more()
        except ConnectionResetError:
            raise _ConnectionDropped()
        lines = self._buf[:end].split(b"\r\n")
        self._buf = self._buf[end + 4:]
        status = int(lines[0].split(None, 2)[1])
        hdrs, length, chunked, close = [], None, False, False
        for ln in lines[1:]:
            k, _, v = ln.partition(b":")
            k, v = k.decode("latin-1"), v.strip().decode("latin-1")
            hdrs.append((k, v))
            lk = k.lower()
            if lk == "content-length":
                length = int(v)
            elif lk == "transfer-encoding" and "chunked" in v.lower():
                chunked = True
            elif lk == "connection" and v.lower() == "close":
                close = True
        if chunked:
            parts = []
            while True:
                while b"\r\n" not in self._buf:
                    self._more()
                size_line, _, self._buf = self._buf.partition(b"\r\n")
                size = int(size_line.split(b";")[0], 16)
                while len(self._buf) < size + 2:
                    self._more()
                parts.append(self._buf[:size])
                self._buf = self._buf[size + 2:]
                if size == 0:
                    break
            payload = b"".join(parts)
        elif length is not None:
            if len(self._buf) < length:
                parts, have = [self._buf], len(self._buf)
                self._buf = b""
                while have < length:
                    chunk = self._sock.recv(min(1 << 20, length - have))
                    if not chunk:
                        raise _ConnectionDropped()
                    parts.append(chunk)
                    have += len(chunk)
                payload = b"".join(parts)
            else:
                payload, self._buf = self._buf[:length], self._buf[length:]
        elif status in (204, 304) or method == "HEAD":
            payload = b""
        else:  # body delimited by the end of the connection
            parts = [self._buf]
            self._buf = b""
            while True:
                chunk = self._sock.recv(1 << 18)
                if not chunk:
                    break
                parts.append(chunk)
            payload, close = b"".join(parts), True
        return status, hdrs, payload, close


class _ConnectionPool:
    """``concurrency`` persistent connections to one host."""

    def __init__(self, host, port, size, connection_timeout, network_timeout, ssl_context):
        self._host, self._port = host, port
        self._connect_timeout = connection_timeout
        self._network_timeout = network_timeout
        self._ssl_context = ssl_context
        self._idle = queue.LifoQueue()
        self._size = max(1, int(size))
        for _ in range(self._size):
            self._idle.put(None)  # created on first use
        self._closed = False

    def _connect(self):
        return _RawConnection(self._host, self._port, self._connect_timeout, self._network_timeout, self._ssl_context)

    def request(self, method, uri, body, headers):
        conn = self._idle.get()
        try:
            for attempt in (0, 1):
                fresh = conn is None
                if fresh:
                    conn = self._connect()
                try:
                    status, hdrs, payload, close = conn.exchange(method, uri, body, headers)
                    if close:
                        conn.close()
                        conn = None
                    return _HttpResponse(status, hdrs, payload)
                except _ConnectionDropped:
                    # a keep-alive connection the server dropped: reconnect once
                    conn.close()
                    conn = None
                    if attempt or fresh:
                        raise ConnectionResetError("connection closed by the server")
        except Exception:
            if conn is not None:
                conn.close()
                conn = None
            raise
        finally:
            self._idle.put(conn)

    def close(self):
        if self._closed:
            return
        self._closed = True
        for _ in range(self._size):
            try:
                conn = self._idle.get_nowait()
            except queue.Empty:
                break
            if conn is not None:
                conn.close()


_device_compression = {"enabled": False}


def set_device_compression(enabled):
    """Compress request bodies with the device deflate encoder (``tb200_deflate_async``)
    instead of host zlib / gzip (reference :1440-1460).  Off by default: the reference's
    behaviour.  When switched on there is no host fallback: without libtb200 and a GPU the
    next compressed request raises.  The stream is a valid zlib / gzip stream but not
    byte-identical to zlib's own output."""
    _device_compression["enabled"] = bool(enabled)


def _compress_on_device(body, algorithm):
    import numpy as np

    from .. import _native
    from ..device import DeviceOps

    ops = DeviceOps(_native.default_context(0))
    host = np.frombuffer(body, dtype=np.uint8)
    src = ops._scratch(max(len(body), 16))
    if len(body):
        ops.h2d(src.ptr, host.ctypes.data, len(body))
    return ops.deflate(src.ptr, len(body), algorithm)  # synchronises: `host` stays alive until then


class InferAsyncRequest:
    """Handle of an in-flight asynchronous inference request.

    Parameters
    ----------
    greenlet : concurrent.futures.Future
        The future that yields the response (the reference holds a gevent
        greenlet here; the name of the argument is kept).
    verbose : bool
        If True generate verbose output.
    """

    def __init__(self, greenlet, verbose=False):
        self._greenlet = greenlet
        self._verbose = verbose

    def get_result(self, block=True, timeout=None):
        """Wait for and return the :py:class:`InferResult`.

        Raises InferenceServerException when the server fails the request or the
        response does not arrive within ``timeout`` seconds.
        """
        try:
            response = self._greenlet.result(timeout=timeout if block else 0)
        except _FutureTimeout:
            raise_error("failed to obtain inference response")
        _raise_if_error(response)
        return InferResult(response, self._verbose)


def _split_url(url):
    """'host:port/base/path' -> (host, port or None, '/base/path' or '')."""
    hostport, _, path = url.partition("/")
    base = ("/" + path).rstrip("/") if path else ""
    if hostport.startswith("["):  # [ipv6]:port
        host, _, rest = hostport[1:].partition("]")
        port = int(rest[1:]) if rest.startswith(":") else None
    else:
        host, _, p = hostport.partition(":")
        port = int(p) if p else None
    return host, port, base


class InferenceServerClient(InferenceServerClientBase):
    """Client of the inference server's HTTP/REST endpoint.  Not thread safe: one
    client object per thread (reference :104-108).

    Parameters
    ----------
    url : str
        ``host:port[/base-path]`` without scheme, e.g. ``localhost:8000``.
    verbose : bool
        Print requests and responses.
    concurrency : int
        Number of connections kept to the server (default 1).
    connection_timeout, network_timeout : float
        Seconds (default 60.0 each).
    max_greenlets : int
        Upper bound of worker threads serving ``async_infer`` (None: one per
        connection).
    ssl, ssl_options, ssl_context_factory, insecure
        HTTPS settings as in the reference.
    """

    def __init__(self, url, verbose=False, concurrency=1, connection_timeout=60.0, network_timeout=60.0,
                 max_greenlets=None, ssl=False, ssl_options=None, ssl_context_factory=None, insecure=False):
        super().__init__()
        if url.startswith("http://") or url.startswith("https://"):
            raise_error("url should not include the scheme")
        host, port, base = _split_url(url)
        context = None
        if ssl:
            context = ssl_context_factory() if ssl_context_factory is not None else _ssl.create_default_context()
            if ssl_options:
                if ssl_options.get("certfile"):
                    context.load_cert_chain(ssl_options["certfile"], ssl_options.get("keyfile"))
                if ssl_options.get("ca_certs"):
                    context.load_verify_locations(ssl_options["ca_certs"])
            if insecure:
                context.check_hostname = False
                context.verify_mode = _ssl.CERT_NONE
        if port is None:
            port = 443 if ssl else 80
        self._base_uri = base
        self._pool_conns = _ConnectionPool(host, port, concurrency, connection_timeout, network_timeout, context)
        workers = max_greenlets if max_greenlets else max(1, int(concurrency))
        self._pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="tb200-http")
        self._verbose = verbose
        self._closed = False

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
        """Close the client; later calls fail."""
        if not getattr(self, "_closed", True):
            self._closed = True
            self._pool.shutdown(wait=True)
            self._pool_conns.close()

    # -- transport ----------------------------------------------------------------
    def _prepare(self, request_uri, headers, query_params):
        request = Request(headers)
        self._call_plugin(request)
        headers = request.headers
        self._validate_headers(headers)
        uri = self._base_uri + "/" + request_uri
        if query_params is not None:
            uri = uri + "?" + _get_query_string(query_params)
        out = {str(k): str(v) for k, v in headers.items()} if headers else {}
        return uri, out

    def _get(self, request_uri, headers, query_params):
        """Issue a GET; returns the response object (reference :228-257)."""
        uri, hdrs = self._prepare(request_uri, headers, query_params)
        if self._verbose:
            print("GET {}, headers {}".format(uri, hdrs))
        response = self._pool_conns.request("GET", uri, None, hdrs)
        if self._verbose:
            print(response)
        return response

    def _post(self, request_uri, request_body, headers, query_params):
        """Issue a POST; returns the response object (reference :259-307)."""
        uri, hdrs = self._prepare(request_uri, headers, query_params)
        if self._verbose:
            print("POST {}, headers {}\n{}".format(uri, hdrs, request_body))
        if isinstance(request_body, str):
            request_body = request_body.encode("utf-8")
        response = self._pool_conns.request("POST", uri, request_body, hdrs)
        if self._verbose:
            print(response)
        return response

    def _validate_headers(self, headers):
        """Reject headers this client cannot honour (reference :309-338)."""
        if not headers:
            return
        lowered = {k.lower(): v for k, v in headers.items()}
        if "transfer-encoding" in lowered:
            raise_error(
                "Unsupported HTTP header: 'Transfer-Encoding' is not "
                "supported in the Python client library. Use raw HTTP "
                "request libraries or the C++ client instead for this "
                "header."
            )

    def _get_json(self, request_uri, headers, query_params):
        response = self._get(request_uri, headers, query_params)
        _raise_if_error(response)
        content = response.read()
        if self._verbose:
            print(content)
        return json.loads(content)

    def _post_json(self, request_uri, body, headers, query_params, parse=True):
        response = self._post(request_uri, body, headers, query_params)
        _raise_if_error(response)
        if not parse:
            return None
        content = response.read()
        if self._verbose:
            print(content)
        return json.loads(content)

    @staticmethod
    def _model_uri(model_name, model_version, suffix=""):
        if type(model_version) != str:
            raise_error("model version must be a string")
        uri = "v2/models/{}".format(quote(model_name))
        if model_version != "":
            uri += "/versions/{}".format(model_version)
        return uri + suffix

    # -- health / metadata ------------------------------------------------------------
    def is_server_live(self, headers=None, query_params=None):
        """True when the server is live (``v2/health/live``)."""
        return self._get("v2/health/live", headers, query_params).status_code == 200

    def is_server_ready(self, headers=None, query_params=None):
        """True when the server is ready (``v2/health/ready``)."""
        return self._get("v2/health/ready", headers, query_params).status_code == 200

    def is_model_ready(self, model_name, model_version="", headers=None, query_params=None):
        """True when the model (version) is ready."""
        uri = self._model_uri(model_name, model_version, "/ready")
        return self._get(uri, headers, query_params).status_code == 200

    def get_server_metadata(self, headers=None, query_params=None):
        """Server metadata dict (``v2``)."""
        return self._get_json("v2", headers, query_params)

    def get_model_metadata(self, model_name, model_version="", headers=None, query_params=None):
        """Model metadata dict."""
        return self._get_json(self._model_uri(model_name, model_version), headers, query_params)

    def get_model_config(self, model_name, model_version="", headers=None, query_params=None):
        """Model configuration dict."""
        return self._get_json(self._model_uri(model_name, model_version, "/config"), headers, query_params)

    def get_model_repository_index(self, headers=None, query_params=None):
        """Index of the model repository."""
        return self._post_json("v2/repository/index", "", headers, query_params)

    def load_model(self, model_name, headers=None, query_params=None, config=None, files=None):
        """Ask the server to load (or reload) a model, optionally with a config
        override and override files (``file:<path>`` -> bytes)."""
        load_request = {}
        if config is not None:
            load_request.setdefault("parameters", {})["config"] = config
        if files is not None:
            for path, content in files.items():
                load_request.setdefault("parameters", {})[path] = base64.b64encode(content).decode("ascii")
        uri = "v2/repository/models/{}/load".format(quote(model_name))
        self._post_json(uri, _dumps(load_request), headers, query_params, parse=False)
        if self._verbose:
            print("Loaded model '{}'".format(model_name))

    def unload_model(self, model_name, headers=None, query_params=None, unload_dependents=False):
        """Ask the server to unload a model."""
        body = _dumps({"parameters": {"unload_dependents": unload_dependents}})
        uri = "v2/repository/models/{}/unload".format(quote(model_name))
        self._post_json(uri, body, headers, query_params, parse=False)
        if self._verbose:
            print("Loaded model '{}'".format(model_name))

    def get_inference_statistics(self, model_name="", model_version="", headers=None, query_params=None):
        """Inference statistics of one model (version) or of all models."""
        if model_name != "":
            uri = self._model_uri(model_name, model_version, "/stats")
        else:
            uri = "v2/models/stats"
        return self._get_json(uri, headers, query_params)

    def update_trace_settings(self, model_name=None, settings={}, headers=None, query_params=None):
        """Update trace settings (of a model, or global); returns the new settings."""
        if (model_name is not None) and (model_name != ""):
            uri = "v2/models/{}/trace/setting".format(quote(model_name))
        else:
            uri = "v2/trace/setting"
        return self._post_json(uri, _dumps(settings), headers, query_params)

    def get_trace_settings(self, model_name=None, headers=None, query_params=None):
        """Trace settings of a model, or the global ones."""
        if (model_name is not None) and (model_name != ""):
            uri = "v2/models/{}/trace/setting".format(quote(model_name))
        else:
            uri = "v2/trace/setting"
        return self._get_json(uri, headers, query_params)

    def update_log_settings(self, settings, headers=None, query_params=None):
        """Update the global log settings; returns the new settings."""
        return self._post_json("v2/logging", _dumps(settings), headers, query_params)

    def get_log_settings(self, headers=None, query_params=None):
        """The global log settings."""
        return self._get_json("v2/logging", headers, query_params)

    # -- shared memory control plane ------------------------------------------------------
    def get_system_shared_memory_status(self, region_name="", headers=None, query_params=None):
        """Status of one / all registered system shared memory regions."""
        if region_name != "":
            uri = "v2/systemsharedmemory/region/{}/status".format(quote(region_name))
        else:
            uri = "v2/systemsharedmemory/status"
        return self._get_json(uri, headers, query_params)

    def register_system_shared_memory(self, name, key, byte_size, offset=0, headers=None, query_params=None):
        """Register a system shared memory region with the server."""
        uri = "v2/systemsharedmemory/region/{}/register".format(quote(name))
        body = _dumps({"key": key, "offset": offset, "byte_size": byte_size})
        self._post_json(uri, body, headers, query_params, parse=False)
        if self._verbose:
            print("Registered system shared memory with name '{}'".format(name))

    def unregister_system_shared_memory(self, name="", headers=None, query_params=None):
        """Unregister one region, or all when ``name`` is empty."""
        if name != "":
            uri = "v2/systemsharedmemory/region/{}/unregister".format(quote(name))
        else:
            uri = "v2/systemsharedmemory/unregister"
        self._post_json(uri, "", headers, query_params, parse=False)
        if self._verbose:
            if name != "":
                print("Unregistered system shared memory with name '{}'".format(name))
            else:
                print("Unregistered all system shared memory regions")

    def get_cuda_shared_memory_status(self, region_name="", headers=None, query_params=None):
        """Status of one / all registered CUDA shared memory regions."""
        if region_name != "":
            uri = "v2/cudasharedmemory/region/{}/status".format(quote(region_name))
        else:
            uri = "v2/cudasharedmemory/status"
        return self._get_json(uri, headers, query_params)

    def register_cuda_shared_memory(self, name, raw_handle, device_id, byte_size, headers=None, query_params=None):
        """Register a CUDA shared memory region; ``raw_handle`` is the base64 IPC
        handle from ``cuda_shared_memory.get_raw_handle`` (reference :1129-1175)."""
        if isinstance(raw_handle, (bytes, bytearray)):
            raw_handle = bytes(raw_handle).decode("ascii")
        uri = "v2/cudasharedmemory/region/{}/register".format(quote(name))
        body = _dumps({"raw_handle": {"b64": raw_handle}, "device_id": device_id, "byte_size": byte_size})
        self._post_json(uri, body, headers, query_params, parse=False)
        if self._verbose:
            print("Registered cuda shared memory with name '{}'".format(name))

    def unregister_cuda_shared_memory(self, name="", headers=None, query_params=None):
        """Unregister one region, or all when ``name`` is empty."""
        if name != "":
            uri = "v2/cudasharedmemory/region/{}/unregister".format(quote(name))
        else:
            uri = "v2/cudasharedmemory/unregister"
        self._post_json(uri, "", headers, query_params, parse=False)
        if self._verbose:
            if name != "":
                print("Unregistered cuda shared memory with name '{}'".format(name))
            else:
                print("Unregistered all cuda shared memory regions")

    # -- inference ------------------------------------------------------------------------
    @staticmethod
    def generate_request_body(inputs, outputs=None, request_id="", sequence_id=0, sequence_start=False,
                              sequence_end=False, priority=0, timeout=None, parameters=None):
        """(request body bytes, json_size or None) without sending anything."""
        return _get_inference_request(
            inputs=inputs, request_id=request_id, outputs=outputs, sequence_id=sequence_id,
            sequence_start=sequence_start, sequence_end=sequence_end, priority=priority,
            timeout=timeout, custom_parameters=parameters,
        )

    @staticmethod
    def parse_response_body(response_body, verbose=False, header_length=None, content_encoding=None):
        """InferResult from raw response bytes."""
        return InferResult.from_response_body(response_body, verbose, header_length, content_encoding)

    def _infer_request(self, model_name, inputs, model_version, outputs, request_id, sequence_id,
                       sequence_start, sequence_end, priority, timeout, headers,
                       request_compression_algorithm, response_compression_algorithm, parameters):
        """(uri, body, headers) of an inference POST (reference :1410-1476)."""
        body, json_size = _get_inference_request(
            inputs=inputs, request_id=request_id, outputs=outputs, sequence_id=sequence_id,
            sequence_start=sequence_start, sequence_end=sequence_end, priority=priority,
            timeout=timeout, custom_parameters=parameters,
        )
        extra = {}
        if request_compression_algorithm in ("gzip", "deflate"):
            extra["Content-Encoding"] = request_compression_algorithm
            if _device_compression["enabled"]:
                body = _compress_on_device(body, request_compression_algorithm)
            elif request_compression_algorithm == "gzip":
                body = gzip.compress(body)
            else:
                body = zlib.compress(body)
        if response_compression_algorithm in ("gzip", "deflate"):
            extra["Accept-Encoding"] = response_compression_algorithm
        if json_size is not None:
            extra["Inference-Header-Content-Length"] = json_size
        if extra:
            if headers is None:
                headers = {}
            headers.update(extra)
        return self._model_uri(model_name, model_version, "/infer"), body, headers

    def infer(self, model_name, inputs, model_version="", outputs=None, request_id="", sequence_id=0,
              sequence_start=False, sequence_end=False, priority=0, timeout=None, headers=None,
              query_params=None, request_compression_algorithm=None,
              response_compression_algorithm=None, parameters=None):
        """Run a synchronous inference; returns :py:class:`InferResult`.

        Raises InferenceServerException if the server fails the request.
        """
        uri, body, headers = self._infer_request(
            model_name, inputs, model_version, outputs, request_id, sequence_id, sequence_start,
            sequence_end, priority, timeout, headers, request_compression_algorithm,
            response_compression_algorithm, parameters,
        )
        response = self._post(uri, body, headers, query_params)
        _raise_if_error(response)
        return InferResult(response, self._verbose)

    def async_infer(self, model_name, inputs, model_version="", outputs=None, request_id="", sequence_id=0,
                    sequence_start=False, sequence_end=False, priority=0, timeout=None, headers=None,
                    query_params=None, request_compression_algorithm=None,
                    response_compression_algorithm=None, parameters=None):
        """Send an inference request without waiting; returns an
        :py:class:`InferAsyncRequest` whose ``get_result()`` yields the result."""
        uri, body, headers = self._infer_request(
            model_name, inputs, model_version, outputs, request_id, sequence_id, sequence_start,
            sequence_end, priority, timeout, headers, request_compression_algorithm,
            response_compression_algorithm, parameters,
        )
        future = self._pool.submit(self._post, uri, body, headers, query_params)
        if self._verbose:
            message = "Sent request"
            if request_id != "":
                message = message + " '{}'".format(request_id)
            print(message)
        return InferAsyncRequest(future, self._verbose)
