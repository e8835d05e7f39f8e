"""asyncio HTTP/REST client.

Drop-in for ``tritonclient.http.aio.InferenceServerClient`` (reference:
src/python/library/tritonclient/http/aio/__init__.py:104-775): same constructor and
coroutine methods over aiohttp; request bodies, URIs and result parsing are shared
with the synchronous client (``client_b200.http``).
"""

import base64
import gzip
import json
import zlib
from urllib.parse import quote

import aiohttp

from ..._client import InferenceServerClientBase
from ..._request import Request
from ...utils import InferenceServerException, raise_error
from .. import InferInput, InferRequestedOutput  # noqa: F401  (re-exported like the reference)
from .._infer_result import InferResult
from .._utils import _dumps, _get_inference_request, _get_query_string


async def _get_error(response):
    """InferenceServerException for a non-200 aiohttp response, else None."""
    if response.status == 200:
        return None
    body = None
    try:
        body = (await response.read()).decode("utf-8")
        payload = json.loads(body) if len(body) else {"error": "client received an empty response from the server."}
        return InferenceServerException(msg=payload["error"], status=str(response.status))
    except Exception as e:
        return InferenceServerException(
            msg=f"an exception occurred in the client while decoding the response: {e}\nresponse: {body}",
            status=str(response.status),
            debug_details=body,
        )


async def _raise_if_error(response):
    error = await _get_error(response)
    if error is not None:
        raise error


class _Body:
    """Adapter giving the fully-read aiohttp body the ``get`` / ``read`` interface
    :py:class:`InferResult` expects."""

    def __init__(self, headers, body):
        self._headers, self._body, self._pos = headers, body, 0

    def get(self, key):
        return self._headers.get(key)

    def read(self, length=-1):
        if length == -1:
            chunk, self._pos = self._body[self._pos:], len(self._body)
        else:
            chunk = self._body[self._pos:self._pos + length]
            self._pos += length
        return chunk


class InferenceServerClient(InferenceServerClientBase):
    """asyncio twin of :py:class:`client_b200.http.InferenceServerClient`; single
    threaded use only.

    Parameters
    ----------
    url : str
        ``host:port[/base-path]`` without scheme.
    verbose : bool
    conn_limit : int
        Maximum simultaneous connections (default 100).
    conn_timeout : float
        Total timeout per request in seconds (default 60.0).
    ssl : bool
    ssl_context : ssl.SSLContext
    """

    def __init__(self, url, verbose=False, conn_limit=100, conn_timeout=60.0, ssl=False, ssl_context=None):
        super().__init__()
        if url.startswith("http://") or url.startswith("https://"):
            raise_error("url should not include the scheme")
        scheme = "https://" if ssl else "http://"
        self._url = scheme + (url if url[-1] != "/" else url[:-1])
        self._conn = aiohttp.TCPConnector(ssl=ssl_context, limit=conn_limit)
        self._stub = aiohttp.ClientSession(
            connector=self._conn, timeout=aiohttp.ClientTimeout(total=conn_timeout), auto_decompress=False
        )
        self._verbose = verbose

    async def __aenter__(self):
        return self

    async def __aexit__(self, type, value, traceback):
        await self.close()

    async def close(self):
        """Close the client session and its connections."""
        await self._stub.close()
        await self._conn.close()

    def _validate_headers(self, headers):
        if not headers:
            return
        if "transfer-encoding" in {k.lower() for k in headers}:
            raise_error(
                "Unsupported HTTP header: 'Transfer-Encoding' is not "
                "supported in the Python client library. Use raw HTTP "
                "request libraries or the C++ client instead for this "
                "header."
            )

    def _fix_header(self, headers):
        """aiohttp wants string header values (reference :213-226)."""
        if headers is None:
            return None
        return {k: str(v) for k, v in headers.items()}

    def _prepare(self, request_uri, headers, query_params):
        request = Request(headers)
        self._call_plugin(request)
        headers = request.headers
        self._validate_headers(headers)
        uri = self._url + "/" + request_uri
        if query_params is not None:
            uri = uri + "?" + _get_query_string(query_params)
        return uri, self._fix_header(headers)

    async def _get(self, request_uri, headers, query_params):
        uri, headers = self._prepare(request_uri, headers, query_params)
        if self._verbose:
            print("GET {}, headers {}".format(uri, headers))
        response = await self._stub.get(url=uri, headers=headers)
        if self._verbose:
            print(response)
        return response

    async def _post(self, request_uri, request_body, headers, query_params):
        uri, headers = self._prepare(request_uri, headers, query_params)
        if self._verbose:
            print("POST {}, headers {}\n{}".format(uri, headers, request_body))
        if isinstance(request_body, str):
            request_body = request_body.encode("utf-8")
        response = await self._stub.post(url=uri, data=request_body, headers=headers)
        if self._verbose:
            print(response)
        return response

    async def _get_json(self, uri, headers, query_params):
        response = await self._get(uri, headers, query_params)
        await _raise_if_error(response)
        content = await response.read()
        if self._verbose:
            print(content)
        return json.loads(content)

    async def _post_json(self, uri, body, headers, query_params, parse=True):
        response = await self._post(uri, body, headers, query_params)
        await _raise_if_error(response)
        if not parse:
            return None
        content = await response.read()
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

    async def is_server_live(self, headers=None, query_params=None):
        return (await self._get("v2/health/live", headers, query_params)).status == 200

    async def is_server_ready(self, headers=None, query_params=None):
        return (await self._get("v2/health/ready", headers, query_params)).status == 200

    async def is_model_ready(self, model_name, model_version="", headers=None, query_params=None):
        return (await self._get(self._model_uri(model_name, model_version, "/ready"), headers, query_params)).status == 200

    async def get_server_metadata(self, headers=None, query_params=None):
        return await self._get_json("v2", headers, query_params)

    async def get_model_metadata(self, model_name, model_version="", headers=None, query_params=None):
        return await self._get_json(self._model_uri(model_name, model_version), headers, query_params)

    async def get_model_config(self, model_name, model_version="", headers=None, query_params=None):
        return await self._get_json(self._model_uri(model_name, model_version, "/config"), headers, query_params)

    async def get_model_repository_index(self, headers=None, query_params=None):
        return await self._post_json("v2/repository/index", "", headers, query_params)

    async def load_model(self, model_name, headers=None, query_params=None, config=None, files=None):
        load_request = {}
        if config is not None:
            load_request.setdefault("parameters", {})["config"] = config
        for path, content in (files or {}).items():
            load_request.setdefault("parameters", {})[path] = base64.b64encode(content).decode("ascii")
        uri = "v2/repository/models/{}/load".format(quote(model_name))
        await self._post_json(uri, _dumps(load_request), headers, query_params, parse=False)
        if self._verbose:
            print("Loaded model '{}'".format(model_name))

    async def unload_model(self, model_name, headers=None, query_params=None, unload_dependents=False):
        body = _dumps({"parameters": {"unload_dependents": unload_dependents}})
        uri = "v2/repository/models/{}/unload".format(quote(model_name))
        await self._post_json(uri, body, headers, query_params, parse=False)
        if self._verbose:
            print("Loaded model '{}'".format(model_name))

    async def get_inference_statistics(self, model_name="", model_version="", headers=None, query_params=None):
        uri = self._model_uri(model_name, model_version, "/stats") if model_name != "" else "v2/models/stats"
        return await self._get_json(uri, headers, query_params)

    async def update_trace_settings(self, model_name=None, settings={}, headers=None, query_params=None):
        uri = "v2/models/{}/trace/setting".format(quote(model_name)) if model_name else "v2/trace/setting"
        return await self._post_json(uri, _dumps(settings), headers, query_params)

    async def get_trace_settings(self, model_name=None, headers=None, query_params=None):
        uri = "v2/models/{}/trace/setting".format(quote(model_name)) if model_name else "v2/trace/setting"
        return await self._get_json(uri, headers, query_params)

    async def update_log_settings(self, settings, headers=None, query_params=None):
        return await self._post_json("v2/logging", _dumps(settings), headers, query_params)

    async def get_log_settings(self, headers=None, query_params=None):
        return await self._get_json("v2/logging", headers, query_params)

    async def get_system_shared_memory_status(self, region_name="", headers=None, query_params=None):
        uri = ("v2/systemsharedmemory/region/{}/status".format(quote(region_name)) if region_name != ""
               else "v2/systemsharedmemory/status")
        return await self._get_json(uri, headers, query_params)

    async def register_system_shared_memory(self, name, key, byte_size, offset=0, headers=None, query_params=None):
        uri = "v2/systemsharedmemory/region/{}/register".format(quote(name))
        body = _dumps({"key": key, "offset": offset, "byte_size": byte_size})
        await self._post_json(uri, body, headers, query_params, parse=False)
        if self._verbose:
            print("Registered system shared memory with name '{}'".format(name))

    async def unregister_system_shared_memory(self, name="", headers=None, query_params=None):
        uri = ("v2/systemsharedmemory/region/{}/unregister".format(quote(name)) if name != ""
               else "v2/systemsharedmemory/unregister")
        await self._post_json(uri, "", headers, query_params, parse=False)

    async def get_cuda_shared_memory_status(self, region_name="", headers=None, query_params=None):
        uri = ("v2/cudasharedmemory/region/{}/status".format(quote(region_name)) if region_name != ""
               else "v2/cudasharedmemory/status")
        return await self._get_json(uri, headers, query_params)

    async def register_cuda_shared_memory(self, name, raw_handle, device_id, byte_size, headers=None, query_params=None):
        if isinstance(raw_handle, (bytes, bytearray)):
            raw_handle = bytes(raw_handle).decode("ascii")
        uri = "v2/cudasharedmemory/region/{}/register".format(quote(name))
        body = _dumps({"raw_handle": {"b64": raw_handle}, "device_id": device_id, "byte_size": byte_size})
        await self._post_json(uri, body, headers, query_params, parse=False)
        if self._verbose:
            print("Registered cuda shared memory with name '{}'".format(name))

    async def unregister_cuda_shared_memory(self, name="", headers=None, query_params=None):
        uri = ("v2/cudasharedmemory/region/{}/unregister".format(quote(name)) if name != ""
               else "v2/cudasharedmemory/unregister")
        await self._post_json(uri, "", headers, query_params, parse=False)

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
        return InferResult.from_response_body(response_body, verbose, header_length, content_encoding)

    async def infer(self, model_name, inputs, model_version="", outputs=None, request_id="", sequence_id=0,
                    sequence_start=False, sequence_end=False, priority=0, timeout=None, headers=None,
                    query_params=None, request_compression_algorithm=None,
                    response_compression_algorithm=None, parameters=None):
        """Run an inference; returns :py:class:`client_b200.http.InferResult`."""
        body, json_size = _get_inference_request(
            inputs=inputs, request_id=request_id, outputs=outputs, sequence_id=sequence_id,
            sequence_start=sequence_start, sequence_end=sequence_end, priority=priority,
            timeout=timeout, custom_parameters=parameters,
        )
        headers = dict(headers) if headers else {}
        if request_compression_algorithm == "gzip":
            headers["Content-Encoding"] = "gzip"
            body = gzip.compress(body)
        elif request_compression_algorithm == "deflate":
            headers["Content-Encoding"] = "deflate"
            body = zlib.compress(body)
        if response_compression_algorithm in ("gzip", "deflate"):
            headers["Accept-Encoding"] = response_compression_algorithm
        if json_size is not None:
            headers["Inference-Header-Content-Length"] = json_size
        response = await self._post(self._model_uri(model_name, model_version, "/infer"), body, headers or None, query_params)
        await _raise_if_error(response)
        payload = await response.read()
        return InferResult(_Body(response.headers, payload), self._verbose)
