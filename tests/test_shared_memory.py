"""The reference's system shared memory unit tests restated against the drop-in
module (reference: src/python/library/tests/test_shared_memory.py:46-179)."""

import numpy as np
import pytest

import client_b200.utils as utils
import client_b200.utils.shared_memory as shm


@pytest.fixture()
def handles():
    hs = []
    yield hs
    for h in hs:
        shm.destroy_shared_memory_region(h)


def test_lifecycle(handles):
    cpu_tensor = np.ones([4, 4], dtype=np.float32)
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 64))
    assert len(shm.mapped_shared_memory_regions()) == 1
    shm.set_shared_memory_region(handles[0], [cpu_tensor])
    shm_tensor = shm.get_contents_as_numpy(handles[0], np.float32, [4, 4])
    assert np.allclose(cpu_tensor, shm_tensor)
    shm.destroy_shared_memory_region(handles.pop(0))
    assert len(shm.mapped_shared_memory_regions()) == 0


def test_invalid_create_shm(handles):
    with pytest.raises(shm.SharedMemoryException, match="unable to create the shared memory region"):
        handles.append(shm.create_shared_memory_region("dummy_data", "/tb200_dummy_data", -1))


def test_set_region_offset(handles):
    large = np.ones([4, 4], dtype=np.float32)
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 64))
    shm.set_shared_memory_region(handles[0], [large])
    small = np.zeros([2, 4], dtype=np.float32)
    shm.set_shared_memory_region(handles[0], [small], offset=32)
    out = shm.get_contents_as_numpy(handles[0], np.float32, [2, 4], offset=32)
    assert np.allclose(small, out)
    assert np.allclose(shm.get_contents_as_numpy(handles[0], np.float32, [2, 4]), 1.0)


def test_set_region_oversize(handles):
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 32))
    with pytest.raises(shm.SharedMemoryException, match="unable to set the shared memory region"):
        shm.set_shared_memory_region(handles[0], [np.ones([4, 4], dtype=np.float32)])


def test_duplicate_key(handles):
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 32))
    with pytest.raises(shm.SharedMemoryException, match="unable to create the shared memory region"):
        handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 32, create_only=True))
    with pytest.warns(UserWarning, match="region size is 32 instead of requested 64"):
        handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 64))
    assert len(shm.mapped_shared_memory_regions()) == 1
    with pytest.raises(shm.SharedMemoryException, match="unable to set the shared memory region"):
        shm.set_shared_memory_region(handles[-1], [np.ones([4, 4], dtype=np.float32)])


def test_destroy_duplicate(handles):
    assert len(shm.mapped_shared_memory_regions()) == 0
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 64))
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 32))
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 32))
    assert len(shm.mapped_shared_memory_regions()) == 1
    shm.destroy_shared_memory_region(handles.pop(0))
    shm.destroy_shared_memory_region(handles.pop(0))
    assert len(shm.mapped_shared_memory_regions()) == 1
    shm.destroy_shared_memory_region(handles.pop(0))
    assert len(shm.mapped_shared_memory_regions()) == 0
    # unlinked: creating with create_only works again
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 16, create_only=True))


def test_numpy_bytes(handles):
    int_tensor = np.arange(start=0, stop=16, dtype=np.int32)
    bytes_tensor = np.array([str(x).encode("utf-8") for x in int_tensor.flatten()], dtype=object)
    bytes_tensor = bytes_tensor.reshape(int_tensor.shape)
    serialized = utils.serialize_byte_tensor(bytes_tensor)
    byte_size = utils.serialized_byte_size(serialized)
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", byte_size))
    shm.set_shared_memory_region(handles[0], [serialized])
    out = shm.get_contents_as_numpy(handles[0], np.object_, [16])
    assert np.array_equal(bytes_tensor, out)


def test_argument_errors(handles):
    handles.append(shm.create_shared_memory_region("shm_name", "tb200_shm_key", 64))
    with pytest.raises(shm.SharedMemoryException, match="list/tuple of numpy arrays"):
        shm.set_shared_memory_region(handles[0], np.zeros(4))
    with pytest.raises(shm.SharedMemoryException, match="each element of input_values must be a numpy array"):
        shm.set_shared_memory_region(handles[0], [[1, 2]])


def test_two_int32_inputs_like_simple_shm_client(handles):
    """simple_http_shm_client.py layout: INPUT0/INPUT1 back to back in one region."""
    a = np.arange(16, dtype=np.int32)
    b = np.ones(16, dtype=np.int32)
    handles.append(shm.create_shared_memory_region("input_data", "/tb200_input_simple", 128))
    shm.set_shared_memory_region(handles[0], [a, b])
    assert np.array_equal(shm.get_contents_as_numpy(handles[0], np.int32, [16], offset=64), b)
