// POSIX shared memory helpers of the reference C++ library (src/c++/library/shm_utils.h:
// CreateSharedMemoryRegion, MapSharedMemory, CloseSharedMemory, UnlinkSharedMemoryRegion,
// UnmapSharedMemory) for code written against `#include "shm_utils.h"`.  Same signatures and
// error texts' meaning; header-only over shm_open / ftruncate / mmap.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>

#include "../tb200_client.h"

namespace tb200 { namespace client {

inline Error CreateSharedMemoryRegion(std::string shm_key, size_t byte_size, int* shm_fd) {
  *shm_fd = shm_open(shm_key.c_str(), O_RDWR | O_CREAT, S_IRUSR | S_IWUSR);
  if (*shm_fd == -1) return Error("unable to get shared memory descriptor for shared-memory key '" + shm_key + "'");
  if (ftruncate(*shm_fd, static_cast<off_t>(byte_size)) == -1) {
    return Error("unable to initialize shared-memory key '" + shm_key + "' to requested size: " + std::to_string(byte_size) + " bytes");
  }
  return Error::Success;
}

inline Error MapSharedMemory(int shm_fd, size_t offset, size_t byte_size, void** shm_addr) {
  *shm_addr = mmap(nullptr, byte_size, PROT_READ | PROT_WRITE, MAP_SHARED, shm_fd, static_cast<off_t>(offset));
  if (*shm_addr == MAP_FAILED) return Error("unable to process address space or shared-memory descriptor: " + std::to_string(shm_fd));
  return Error::Success;
}

inline Error CloseSharedMemory(int shm_fd) {
  if (close(shm_fd) == -1) return Error("unable to close shared-memory descriptor: " + std::to_string(shm_fd));
  return Error::Success;
}

inline Error UnlinkSharedMemoryRegion(std::string shm_key) {
  if (shm_unlink(shm_key.c_str()) == -1) return Error("unable to unlink shared memory for key '" + shm_key + "'");
  return Error::Success;
}

inline Error UnmapSharedMemory(void* shm_addr, size_t byte_size) {
  if (munmap(shm_addr, byte_size) == -1) return Error("unable to munmap shared memory region");
  return Error::Success;
}

}}  // namespace tb200::client

namespace triton { namespace client = ::tb200::client; }
