"""DLPack view over a shared memory region (system or CUDA).

Drop-in for ``tritonclient.utils._shared_memory_tensor.SharedMemoryTensor``
(reference: src/python/library/tritonclient/utils/_shared_memory_tensor.py:32-87).
The object does not own the region: it is invalid once the region is modified
or destroyed.
"""

import ctypes
from typing import Any, Iterable

from . import _dlpack


class SharedMemoryTensor:
    def __init__(
        self,
        dtype: str,
        shape: Iterable,
        shm_addr: Any,
        offset: int,
        byte_size: int,
        device_id: int,
    ) -> None:
        self._dtype = dtype
        self._shape = list(shape)
        self._shm_addr = shm_addr
        self._offset = offset
        self._byte_size = byte_size
        self._device_id = device_id
        if device_id != -1:
            self._dl_device = (_dlpack.DLDeviceType.kDLCUDA, device_id)
        else:
            self._dl_device = (_dlpack.DLDeviceType.kDLCPU, 0)

    def __dlpack__(self, stream=None, **_unused):
        ctx = _dlpack.DataViewContext(self._shape)
        raw = _dlpack._api.PyMem_RawMalloc(ctypes.sizeof(_dlpack.DLManagedTensor))
        managed = _dlpack.DLManagedTensor.from_address(raw)
        t = managed.dl_tensor
        t.data = int(self._shm_addr)
        t.device = _dlpack.DLDevice(*self._dl_device)
        t.ndim = len(self._shape)
        t.dtype = _dlpack.triton_to_dlpack_dtype(self._dtype)
        t.shape = ctx._shape
        t.strides = ctx._strides
        t.byte_offset = self._offset
        managed.manager_ctx = ctx.as_manager_ctx()
        managed.deleter = _dlpack.managed_tensor_deleter
        return _dlpack._api.PyCapsule_New(raw, _dlpack.c_str_dltensor, _dlpack.pycapsule_deleter)

    def __dlpack_device__(self):
        return self._dl_device
