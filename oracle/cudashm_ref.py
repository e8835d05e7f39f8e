"""CPU/driver-side restatement of the reference's cuda_shared_memory data path with
cuda-python, for timing comparisons on the GPU box.  TEST / BENCH INFRASTRUCTURE.

Restates src/python/library/tritonclient/utils/cuda_shared_memory/__init__.py:
  :128-137  create (cudaSetDevice, cudaMalloc, cudaIpcGetMemHandle)
  :203-231  set_shared_memory_region (np.ascontiguousarray().flatten(), cudaMemcpyAsync
            from the pageable array, cudaStreamSynchronize)
  :262-304  get_contents_as_numpy (D2H of the WHOLE region into a ctypes buffer, sync,
            np.frombuffer + np.copy)
"""

import ctypes

import numpy as np


class RefRegion:
    def __init__(self, byte_size, device_id=0):
        import cuda.bindings.runtime as cudart

        self.rt = cudart
        (err,) = cudart.cudaSetDevice(device_id)
        assert err == cudart.cudaError_t.cudaSuccess, err
        err, self.ptr = cudart.cudaMalloc(byte_size)
        assert err == cudart.cudaError_t.cudaSuccess, err
        err, self.handle = cudart.cudaIpcGetMemHandle(self.ptr)
        assert err == cudart.cudaError_t.cudaSuccess, err
        err, self.stream = cudart.cudaStreamCreate()
        assert err == cudart.cudaError_t.cudaSuccess, err
        self.byte_size = byte_size

    def set(self, arrays):
        rt = self.rt
        off = 0
        for a in arrays:
            a = np.ascontiguousarray(a).flatten()
            n = a.size * a.itemsize
            (err,) = rt.cudaMemcpyAsync(self.ptr + off, a.ctypes.data, n, rt.cudaMemcpyKind.cudaMemcpyDefault, self.stream)
            assert err == rt.cudaError_t.cudaSuccess, err
            off += n
        (err,) = rt.cudaStreamSynchronize(self.stream)
        assert err == rt.cudaError_t.cudaSuccess, err

    def get(self, dtype, shape):
        rt = self.rt
        host = (ctypes.c_char * self.byte_size)()
        (err,) = rt.cudaMemcpyAsync(host, self.ptr, self.byte_size, rt.cudaMemcpyKind.cudaMemcpyDefault, self.stream)
        assert err == rt.cudaError_t.cudaSuccess, err
        (err,) = rt.cudaStreamSynchronize(self.stream)
        assert err == rt.cudaError_t.cudaSuccess, err
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        buf = ctypes.cast(host, ctypes.POINTER(ctypes.c_byte * n))[0]
        return np.reshape(np.copy(np.frombuffer(buf, dtype=dtype)), shape)

    def close(self):
        self.rt.cudaFree(self.ptr)
        self.rt.cudaStreamDestroy(self.stream)
