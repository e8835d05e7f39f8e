"""Bidirectional inference stream plumbing.

Drop-in for ``tritonclient.grpc._infer_stream`` (reference:
src/python/library/tritonclient/grpc/_infer_stream.py:39-191): a request queue
feeding the gRPC request iterator and a reader thread that runs the user
callback for every response.
"""

import queue
import threading

import grpc

from ..utils import InferenceServerException, raise_error
from ._infer_result import InferResult
from ._utils import get_cancelled_error, get_error_grpc


class _InferStream:
    """One active ``ModelStreamInfer`` call.

    Parameters
    ----------
    callback : callable(result, error)
        Invoked on the reader thread for every response; ``error`` is None on
        success.
    verbose : bool
        Print stream events.
    """

    def __init__(self, callback, verbose):
        self._callback = callback
        self._verbose = verbose
        self._request_queue = queue.Queue()
        self._handler = None
        self._cancelled = False
        self._active = True
        self._response_iterator = None

    def __del__(self):
        self.close(cancel_requests=True)

    def close(self, cancel_requests=False):
        """Close the stream: cancel pending requests, or (default) drain them first."""
        if cancel_requests and self._response_iterator:
            self._response_iterator.cancel()
            self._cancelled = True
        if self._handler is not None:
            if not self._cancelled:
                self._request_queue.put(None)  # ends the request iterator
            if self._handler.is_alive():
                self._handler.join()
                if self._verbose:
                    print("stream stopped...")
            self._handler = None

    def _init_handler(self, response_iterator):
        """Start the reader thread over the response iterator."""
        self._response_iterator = response_iterator
        if self._handler is not None:
            raise_error("Attempted to initialize already initialized InferStream")
        self._handler = threading.Thread(target=self._process_response)
        self._handler.start()
        if self._verbose:
            print("stream started...")

    def _enqueue_request(self, request):
        """Queue a ModelInferRequest for the request iterator."""
        if not self._active:
            raise_error(
                "The stream is no longer in valid state, the error detail "
                "is reported through provided callback. A new stream should "
                "be started after stopping the current stream."
            )
        self._request_queue.put(request)

    def _get_request(self):
        """Next queued request (blocks); None ends the stream."""
        return self._request_queue.get()

    def _process_response(self):
        """Reader thread: response -> InferResult / error -> callback."""
        try:
            for response in self._response_iterator:
                if self._verbose:
                    print(response)
                if response.error_message != "":
                    self._callback(result=None, error=InferenceServerException(msg=response.error_message))
                else:
                    self._callback(result=InferResult(response.infer_response), error=None)
        except grpc.RpcError as rpc_error:
            # the stream died: remember whether it can still be used and report
            self._active = self._response_iterator.is_active()
            if rpc_error.code() == grpc.StatusCode.CANCELLED:
                error = get_cancelled_error(rpc_error.details())
            else:
                error = get_error_grpc(rpc_error)
            self._callback(result=None, error=error)


class _RequestIterator:
    """Iterator handed to gRPC as the request stream."""

    def __init__(self, stream):
        self._stream = stream

    def __iter__(self):
        return self

    def __next__(self):
        request = self._stream._get_request()
        if request is None:
            raise StopIteration
        return request
