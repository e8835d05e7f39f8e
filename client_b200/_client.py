"""Plugin registry shared by the HTTP and gRPC clients (reference: PY/_client.py:31-84)."""

from .utils import raise_error


class InferenceServerClientBase:
    def __init__(self):
        self._plugin = None

    def _call_plugin(self, request):
        if self._plugin is not None:
            self._plugin(request)

    def register_plugin(self, plugin):
        if self._plugin is not None:
            raise_error(
                "A plugin is already registered. Please "
                "unregister the previous plugin first before"
                " registering a new plugin."
            )
        self._plugin = plugin

    def plugin(self):
        return self._plugin

    def unregister_plugin(self):
        if self._plugin is None:
            raise_error("No plugin has been registered.")
        self._plugin = None
