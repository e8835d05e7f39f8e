"""Client plugin interface (reference: PY/_plugin.py:31-48)."""

from abc import ABC, abstractmethod


class InferenceServerClientPlugin(ABC):
    """Base class of client plugins: ``__call__(request)`` is invoked before every
    request and may modify ``request.headers`` in place."""

    @abstractmethod
    def __call__(self, request):
        pass
