"""``tritonclient.grpc.auth`` (reference: src/python/library/tritonclient/grpc/auth/__init__.py):
re-exports the basic-auth header plugin of ``client_b200._auth``."""

from ..._auth import BasicAuth

__all__ = ["BasicAuth"]
