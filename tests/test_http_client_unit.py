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
