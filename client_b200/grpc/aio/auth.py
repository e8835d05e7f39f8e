from ..._auth import BasicAuth

__all__ = ["BasicAuth"]
