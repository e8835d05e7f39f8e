"""Handle types of the CUDA shared memory module.

Drop-in for ``tritonclient.utils.cuda_shared_memory._utils`` (reference:
src/python/library/tritonclient/utils/cuda_shared_memory/_utils.py:49-128);
the CUDA runtime calls of the reference (cudaMalloc / cudaIpcGetMemHandle /
cudaFree / stream create) live behind the C ABI of libtb200 (include/tb200.h).
"""

import ctypes
from typing import Any

from ... import _native


class CudaSharedMemoryException(Exception):
    """Exception indicating non-Success status (reference :49-64)."""

    def __init__(self, msg):
        self._msg = msg

    def __str__(self):
        return super().__str__() if self._msg is None else self._msg


class IpcMemHandle:
    """The 64-byte ``cudaIpcMemHandle_t``; ``reserved`` matches the attribute the
    reference reads from cuda-python's handle object (:170)."""

    __slots__ = ("reserved",)

    def __init__(self, raw: bytes) -> None:
        self.reserved = bytes(raw)


class CudaSharedMemoryRegion:
    """A device allocation exported over CUDA IPC.  Attribute names follow the
    reference (:76-80); the allocation is released when the object dies
    (:88-100)."""

    def __init__(
        self,
        triton_shm_name: str,
        cuda_shm_handle: IpcMemHandle,
        base_addr: Any,
        byte_size: int,
        device_id: int,
        native_region=None,
    ) -> None:
        self._triton_shm_name = triton_shm_name
        self._cuda_shm_handle = cuda_shm_handle
        self._base_addr = base_addr
        self._byte_size = byte_size
        self._device_id = device_id
        self._native = native_region  # tb200_region*

    def __del__(self):
        native = getattr(self, "_native", None)
        if native:
            try:
                _native.load().tb200_region_destroy(native)
            except Exception:
                pass
            self._native = None


class CudaStream:
    """Per-device stream object kept for API compatibility (reference
    :103-121): ``_stream`` is the cudaStream_t of the device's tb200 context."""

    def __init__(self, device_id):
        self._ctx = _native.default_context(device_id)
        self._stream = ctypes.c_void_p(_native.load().tb200_ctx_stream(self._ctx.handle)).value or 0

    def getPtr(self):
        return self._stream


def maybe_set_device(device_id):
    """No-op: libtb200 switches and restores the current device inside every
    call (reference :124-128 restores it from Python)."""
    return None
