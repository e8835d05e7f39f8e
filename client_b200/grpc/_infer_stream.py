"""Bidirectional inference stream plumbing.

Drop-in for ``tritonclient.grpc._infer_stream`` (reference:
src/python/library/tritonclient/grpc/_infer_stream.py:39-191).  The names the client uses
(``_InferStream`` with ``_init_handler`` / ``_enqueue_request`` / ``close``, ``_RequestIterator``)
are kept; inside, outgoing requests sit in a deque guarded by a condition variable and are
handed to gRPC by a generator, and one reader thread turns responses into callbacks.
"""

import collections
import threading

import grpc

from ..utils import InferenceServerException, raise_error
from ._infer_result import InferResult
from ._utils import get_cancelled_error, get_error_grpc

_CLOSED = object()  # sentinel: no more requests

_DEAD_STREAM = (
    "The stream is no longer in valid state, the error detail "
    "is reported through provided callback. A new stream should "
    "be started after stopping the current stream."
)


class _InferStream:
    """One active ``ModelStreamInfer`` call: ``callback(result, error)`` runs on the reader
    thread for every response (``error`` is None on success)."""

    def __init__(self, callback, verbose):
        self._callback, self._verbose = callback, verbose
        self._outbox = collections.deque()
        self._wake = threading.Condition()
        self._responses = None  # the gRPC call object, also the response iterator
        self._reader = None
        self._usable = True

    def __del__(self):
        self.close(cancel_requests=True)

    # ---- requests --------------------------------------------------------------------
    def _enqueue_request(self, request):
        if not self._usable:
            raise_error(_DEAD_STREAM)
        with self._wake:
            self._outbox.append(request)
            self._wake.notify()

    def _outgoing(self):
        """Generator gRPC pulls the request stream from."""
        while True:
            with self._wake:
                while not self._outbox:
                    self._wake.wait()
                item = self._outbox.popleft()
            if item is _CLOSED:
                return
            yield item

    # ---- responses ---------------------------------------------------------------------
    def _init_handler(self, response_iterator):
        if self._reader is not None:
            raise_error("Attempted to initialize already initialized InferStream")
        self._responses = response_iterator
        self._reader = threading.Thread(target=self._pump, name="tb200-grpc-stream")
        self._reader.start()
        if self._verbose:
            print("stream started...")

    def _pump(self):
        deliver = self._callback
        try:
            for message in self._responses:
                if self._verbose:
                    print(message)
                if message.error_message:
                    deliver(result=None, error=InferenceServerException(msg=message.error_message))
                else:
                    deliver(result=InferResult(message.infer_response), error=None)
        except grpc.RpcError as failure:
            # the call ended abnormally: later enqueues must fail, the user hears about it once
            self._usable = self._responses.is_active()
            cancelled = failure.code() == grpc.StatusCode.CANCELLED
            deliver(result=None, error=get_cancelled_error(failure.details()) if cancelled else get_error_grpc(failure))

    def close(self, cancel_requests=False):
        """Stop the stream.  Default: let queued requests go out and wait for their responses;
        ``cancel_requests=True`` cancels the call right away."""
        if cancel_requests and self._responses is not None:
            self._responses.cancel()
        reader, self._reader = self._reader, None
        if reader is None:
            return
        with self._wake:  # also after a cancel: releases gRPC's request-consuming thread
            self._outbox.append(_CLOSED)
            self._wake.notify()
        if reader.is_alive():
            reader.join()
            if self._verbose:
                print("stream stopped...")


class _RequestIterator:
    """What ``ModelStreamInfer`` receives as its request iterator."""

    def __init__(self, stream):
        self._pull = stream._outgoing()

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._pull)
