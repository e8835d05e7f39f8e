"""POSIX (system) shared memory regions.

Drop-in for ``tritonclient.utils.shared_memory`` (reference:
src/python/library/tritonclient/utils/shared_memory/__init__.py:50-257).  Same
functions, arguments, error texts and ref-counted unlink semantics.  The mapping
is done with ``shm_open`` + ``mmap`` directly instead of
``multiprocessing.shared_memory`` so that a process that merely attaches to a
region (the server side of the loop) never unlinks it through Python's resource
tracker.
"""

import mmap
import os
import warnings

import _posixshmem
import numpy as np

# key -> {"needs_unlink": bool, "active_handle_count": int}   (reference :36)
_key_mapping = {}


class SharedMemoryException(Exception):
    """Exception type for shared memory related error (reference :254-257)."""

    pass


class _PosixRegion:
    """One mapping of a POSIX shared memory object; exposes the attributes the
    reference reads from ``multiprocessing.shared_memory.SharedMemory``
    (``buf``, ``size``, ``name``, ``close``, ``unlink``)."""

    def __init__(self, key, create=False, size=0):
        self._name = key if key.startswith("/") else "/" + key
        flags = os.O_RDWR
        if create:
            if not isinstance(size, int) or size <= 0:
                raise ValueError("'size' must be a positive integer")
            flags |= os.O_CREAT | os.O_EXCL
        fd = _posixshmem.shm_open(self._name, flags, mode=0o600)
        try:
            if create:
                os.ftruncate(fd, size)
            self.size = os.fstat(fd).st_size
            self._mmap = mmap.mmap(fd, self.size)
        except Exception:
            os.close(fd)
            if create:
                try:
                    _posixshmem.shm_unlink(self._name)
                except OSError:
                    pass
            raise
        os.close(fd)
        self.buf = memoryview(self._mmap)

    @property
    def name(self):
        return self._name[1:]

    def close(self):
        if self.buf is not None:
            try:
                self.buf.release()
            except BufferError:
                pass
            self.buf = None
        if self._mmap is not None:
            try:
                self._mmap.close()
            except BufferError:
                # numpy views handed out by get_contents_as_numpy are still alive;
                # the mapping goes away with them
                pass
            self._mmap = None

    def unlink(self):
        _posixshmem.shm_unlink(self._name)


class SharedMemoryRegion:
    def __init__(self, triton_shm_name: str, shm_key: str) -> None:
        self._triton_shm_name = triton_shm_name
        self._shm_key = shm_key
        self._mpsm_handle = None


def _track(shm_key, created):
    entry = _key_mapping.setdefault(shm_key, {"needs_unlink": False, "active_handle_count": 0})
    if created:
        entry["needs_unlink"] = True
    entry["active_handle_count"] += 1


def create_shared_memory_region(triton_shm_name, shm_key, byte_size, create_only=False):
    """Return a handle of the system shared memory region with the given key,
    creating it when it does not exist (reference :50-112).

    Raises
    ------
    SharedMemoryException
        If unable to create the shared memory region.
    """
    shm_handle = SharedMemoryRegion(triton_shm_name, shm_key)
    if not create_only:
        try:
            shm_handle._mpsm_handle = _PosixRegion(shm_key)
            _track(shm_key, created=False)
        except FileNotFoundError:
            pass  # not there yet: create it below
    if shm_handle._mpsm_handle is None:
        try:
            shm_handle._mpsm_handle = _PosixRegion(shm_key, create=True, size=byte_size)
        except Exception as ex:
            raise SharedMemoryException("unable to create the shared memory region") from ex
        _track(shm_key, created=True)

    if byte_size > shm_handle._mpsm_handle.size:
        warnings.warn(
            f"reusing shared memory region with key '{shm_key}', region size is "
            f"{shm_handle._mpsm_handle.size} instead of requested {byte_size}"
        )
    return shm_handle


def set_shared_memory_region(shm_handle, input_values, offset=0):
    """Copy numpy arrays back to back into the region starting at ``offset``
    (reference :115-163).

    Raises
    ------
    SharedMemoryException
        If unable to set values in the system shared memory region.
    """
    if not isinstance(input_values, (list, tuple)):
        raise SharedMemoryException("input_values must be specified as a list/tuple of numpy arrays")
    for input_value in input_values:
        if not isinstance(input_value, np.ndarray):
            raise SharedMemoryException("each element of input_values must be a numpy array")

    try:
        buf = shm_handle._mpsm_handle.buf
        for input_value in input_values:
            if input_value.dtype == np.object_:
                # a serialised BYTES tensor: 0-d object array holding the bytes
                payload = input_value.item()
                n = len(payload)
                if offset + n > len(buf):
                    raise ValueError("tensor does not fit the region")
                buf[offset : offset + n] = payload
                offset += n
            else:
                n = input_value.nbytes
                view = np.ndarray(input_value.shape, input_value.dtype, buffer=buf[offset:])
                view[...] = input_value
                offset += n
    except Exception as ex:
        raise SharedMemoryException("unable to set the shared memory region") from ex


def get_contents_as_numpy(shm_handle, datatype, shape, offset=0):
    """numpy array over (fixed-size types: zero copy) or decoded from (BYTES)
    the region contents (reference :166-210)."""
    buf = shm_handle._mpsm_handle.buf
    if (datatype != np.object_) and (datatype != np.bytes_):
        return np.ndarray(shape, datatype, buffer=buf[offset:])
    count = int(np.prod(shape))
    strs = []
    pos = offset
    # the reference reads at least one element even for an empty shape (:196)
    for _ in range(max(count, 1)):
        n = int.from_bytes(buf[pos : pos + 4], "little")
        pos += 4
        strs.append(bytes(buf[pos : pos + n]))
        pos += n
    return np.reshape(np.array(strs, dtype=object), shape)


def mapped_shared_memory_regions():
    """Keys of all regions mapped by this process and not yet destroyed
    (reference :213-222)."""
    return list(_key_mapping.keys())


def destroy_shared_memory_region(shm_handle):
    """Release the handle; unlink the object when the last handle that this
    process created for the key goes away (reference :225-251).

    Raises
    ------
    SharedMemoryException
        If unable to unlink the shared memory region.
    """
    shm_handle._mpsm_handle.close()
    entry = _key_mapping[shm_handle._shm_key]
    entry["active_handle_count"] -= 1
    if entry["active_handle_count"] == 0:
        try:
            if entry["needs_unlink"]:
                shm_handle._mpsm_handle.unlink()
        except OSError as ex:
            raise SharedMemoryException("unable to unlink the shared memory region") from ex
        finally:
            _key_mapping.pop(shm_handle._shm_key)
