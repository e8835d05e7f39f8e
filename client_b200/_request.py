"""Request object handed to plugins (reference: PY/_request.py:29-39)."""


class Request:
    def __init__(self, headers):
        self.headers = headers if headers is not None else {}
