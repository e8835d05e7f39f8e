"""DLPack structures (ctypes) and helpers for exchanging tensors with shared
memory regions.

Drop-in for ``tritonclient.utils._dlpack`` (reference:
src/python/library/tritonclient/utils/_dlpack.py).  The struct layouts follow
dlpack.h (DLDevice, DLDataType, DLTensor, DLManagedTensor); names and helper
signatures follow the reference so that code written against it keeps working.
"""

import ctypes

_api = ctypes.pythonapi
_api.PyMem_RawMalloc.restype = ctypes.c_void_p
_api.PyMem_RawMalloc.argtypes = [ctypes.c_size_t]
_api.PyMem_RawFree.restype = None
_api.PyMem_RawFree.argtypes = [ctypes.c_void_p]
_api.PyCapsule_New.restype = ctypes.py_object
_api.PyCapsule_New.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
_api.PyCapsule_GetPointer.restype = ctypes.c_void_p
_api.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_api.PyCapsule_IsValid.restype = ctypes.c_int
_api.PyCapsule_IsValid.argtypes = [ctypes.py_object, ctypes.c_char_p]
_api.PyCapsule_SetDestructor.restype = ctypes.c_int
_api.PyCapsule_SetDestructor.argtypes = [ctypes.py_object, ctypes.c_void_p]

c_str_dltensor = b"dltensor"


class DLDeviceType(ctypes.c_int):
    """dlpack.h DLDeviceType (reference :57-71)."""

    kDLCPU = 1
    kDLCUDA = 2
    kDLCUDAHost = 3
    kDLOpenCL = 4
    kDLVulkan = 7
    kDLMetal = 8
    kDLVPI = 9
    kDLROCM = 10
    kDLROCMHost = 11
    kDLExtDev = 12
    kDLCUDAManaged = 13
    kDLOneAPI = 14
    kDLWebGPU = 15
    kDLHexagon = 16


class DLDevice(ctypes.Structure):
    _fields_ = [("device_type", DLDeviceType), ("device_id", ctypes.c_int)]


class DLDataTypeCode(ctypes.c_uint8):
    kDLInt = 0
    kDLUInt = 1
    kDLFloat = 2
    kDLOpaquePointer = 3
    kDLBfloat = 4
    kDLComplex = 5
    kDLBool = 6


class DLDataType(ctypes.Structure):
    _fields_ = [
        ("type_code", DLDataTypeCode),
        ("bits", ctypes.c_uint8),
        ("lanes", ctypes.c_uint16),
    ]


class DLTensor(ctypes.Structure):
    _fields_ = [
        ("data", ctypes.c_void_p),
        ("device", DLDevice),
        ("ndim", ctypes.c_int),
        ("dtype", DLDataType),
        ("shape", ctypes.POINTER(ctypes.c_int64)),
        ("strides", ctypes.POINTER(ctypes.c_int64)),
        ("byte_offset", ctypes.c_uint64),
    ]


_DELETER = ctypes.CFUNCTYPE(None, ctypes.c_void_p)


class DLManagedTensor(ctypes.Structure):
    _fields_ = [
        ("dl_tensor", DLTensor),
        ("manager_ctx", ctypes.c_void_p),
        ("deleter", _DELETER),
    ]


def _raise_error(msg):
    raise Exception(msg) from None


# Views exported over a region do not own the bytes; the manager context only
# keeps the shape array alive until the consumer calls the deleter.
_live_views = {}


class DataViewContext:
    """Keeps the ctypes shape array of an exported view alive (reference :131-146)."""

    def __init__(self, shape) -> None:
        dims = [int(d) for d in shape]
        self._shape = (ctypes.c_int64 * len(dims))(*dims)
        self._strides = ctypes.POINTER(ctypes.c_int64)()  # NULL: compact row-major

    def as_manager_ctx(self) -> ctypes.c_void_p:
        key = id(self)
        _live_views[key] = self
        return ctypes.c_void_p(key)


@_DELETER
def managed_tensor_deleter(handle) -> None:
    managed = DLManagedTensor.from_address(handle)
    _live_views.pop(managed.manager_ctx, None)
    _api.PyMem_RawFree(handle)


@ctypes.CFUNCTYPE(None, ctypes.c_void_p)
def pycapsule_deleter(handle) -> None:
    capsule = ctypes.cast(handle, ctypes.py_object)
    # a consumer renames the capsule to "used_dltensor" and takes ownership
    if _api.PyCapsule_IsValid(capsule, c_str_dltensor):
        managed_tensor_deleter(_api.PyCapsule_GetPointer(capsule, c_str_dltensor))
        _api.PyCapsule_SetDestructor(capsule, None)


_TRITON_TO_DL = {
    "BOOL": (DLDataTypeCode.kDLBool, 8),
    "INT8": (DLDataTypeCode.kDLInt, 8),
    "INT16": (DLDataTypeCode.kDLInt, 16),
    "INT32": (DLDataTypeCode.kDLInt, 32),
    "INT64": (DLDataTypeCode.kDLInt, 64),
    "UINT8": (DLDataTypeCode.kDLUInt, 8),
    "UINT16": (DLDataTypeCode.kDLUInt, 16),
    "UINT32": (DLDataTypeCode.kDLUInt, 32),
    "UINT64": (DLDataTypeCode.kDLUInt, 64),
    "FP16": (DLDataTypeCode.kDLFloat, 16),
    "FP32": (DLDataTypeCode.kDLFloat, 32),
    "FP64": (DLDataTypeCode.kDLFloat, 64),
    "BF16": (DLDataTypeCode.kDLBfloat, 16),
}


def triton_to_dlpack_dtype(dtype):
    """Triton datatype name -> DLDataType (reference :170-216).

    BOOL is exported with 8 bits -- the storage size consumers such as torch
    require (the reference writes bits=1, which torch rejects).
    """
    if dtype == "BYTES":
        _raise_error("DLPack currently doesn't suppose BYTES type")
    entry = _TRITON_TO_DL.get(dtype)
    if entry is None:
        _raise_error("Can not covert unknown data type '{}' to DLPack data type".format(dtype))
    return DLDataType(entry[0], entry[1], 1)


def is_contiguous_data(ndim, shape, stride):
    """True when the strides describe compact C order (reference :219-233)."""
    if (stride is None) or (not bool(stride)):
        return True
    expected = 1
    for i in range(int(ndim) - 1, -1, -1):
        if shape[i] != 1 and stride[i] != expected:
            return False
        expected *= shape[i]
    return True


def get_byte_size(dtype, ndim, shape):
    """Bytes of a compact tensor (reference :236-242)."""
    total = dtype.bits * dtype.lanes // 8
    for i in range(int(ndim)):
        total *= shape[i]
    return total


def get_dlpack_capsule(dlpack_obj, stream=None):
    """PyCapsule of a DLPack producer (reference :245-262)."""
    if hasattr(dlpack_obj, "__dlpack__"):
        if not hasattr(dlpack_obj, "__dlpack_device__"):
            _raise_error("DLPack expects '__dlpack_device__' if '__dlpack__' has been defined")
        device = dlpack_obj.__dlpack_device__()
        if int(device[0]) != DLDeviceType.kDLCUDA:
            return dlpack_obj.__dlpack__()
        return dlpack_obj.__dlpack__(stream=stream)
    return dlpack_obj  # legacy: the capsule itself


def get_dlpack_device(dlpack_obj):
    if hasattr(dlpack_obj, "__dlpack_device__"):
        return dlpack_obj.__dlpack_device__()
    return None


def get_managed_tensor(dlcapsule):
    return DLManagedTensor.from_address(_api.PyCapsule_GetPointer(dlcapsule, c_str_dltensor))
