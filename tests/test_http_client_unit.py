"""The reference's HTTP client unit tests restated against the drop-in
(src/python/library/tests/test_inference_server_client.py:47-117): a non-200 response raises
InferenceServerException for a JSON error body and for a plain-text one (never a JSON decode
error), and patched _get/_post successes pass through."""

from unittest.mock import MagicMock, patch

import pytest

from client_b200.http import InferenceServerClient
from client_b200.http._client import _HttpResponse
from client_b200.http._utils import _raise_if_error
from client_b200.utils import InferenceServerException

JSON_ERROR = b"""{
                  "error":"foo",
                  "status_code":"404"
                  }"""


def test_get_method_success():
    with patch("client_b200.http.InferenceServerClient._get", MagicMock(return_value={"status_code": 200})):
        client = InferenceServerClient("dummy_url")
        assert client._get("dummy_url", None, None)["status_code"] == 200


def test_post_method_success():
    with patch("client_b200.http.InferenceServerClient._post", MagicMock(return_value={"status_code": 200})):
        client = InferenceServerClient("dummy_url")
        assert client._post("dummy_url", "dummy_body", None, None)["status_code"] == 200


def test_get_method_failure():
    with pytest.raises(InferenceServerException) as info:
        _raise_if_error(_HttpResponse(400, [], JSON_ERROR))
    assert info.value.message() == "foo" and info.value.status() == "400"


def test_error_plain_text():
    with pytest.raises(InferenceServerException) as info:
        _raise_if_error(_HttpResponse(404, [], b"error_string"))
    assert "error_string" in info.value.message() and info.value.status() == "404"


def test_success_does_not_raise():
    _raise_if_error(_HttpResponse(200, [("Content-Length", "2")], b"{}"))


def test_one_accept_encoding_header_per_request():
    """The default `Accept-Encoding: identity` gives way to the caller's own (reference: only the requested
    encoding is sent, http/_client.py:1452-1456 of the reference)."""
    import socket
    import threading

    from client_b200.http._client import _RawConnection

    srv = socket.socket()
    srv.bind(("127.0.0.1", 0))
    srv.listen(1)
    seen = []

    def serve():
        c, _ = srv.accept()
        for _ in range(2):
            buf = b""
            while b"\r\n\r\n" not in buf:
                buf += c.recv(65536)
            seen.append(buf)
            c.sendall(b"HTTP/1.1 200 OK\r\nContent-Length: 0\r\n\r\n")
        c.close()

    th = threading.Thread(target=serve)
    th.start()
    conn = _RawConnection("127.0.0.1", srv.getsockname()[1], 5.0, 5.0, None)
    conn.exchange("GET", "/v2/health/live", None, {})
    conn.exchange("GET", "/v2/health/live", None, {"accept-encoding": "gzip"})
    conn.close()
    th.join()
    srv.close()
    assert seen[0].lower().count(b"accept-encoding") == 1 and b"identity" in seen[0]
    assert seen[1].lower().count(b"accept-encoding") == 1 and b"gzip" in seen[1] and b"identity" not in seen[1]
