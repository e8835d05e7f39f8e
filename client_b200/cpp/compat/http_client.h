// Source-compatibility shim: code written against the reference's
// `#include "http_client.h"` / `namespace tc = triton::client;` builds against tb200_client.h.
#pragma once
// the reference headers pull in the CUDA runtime API when built with GPU support; code written
// against them uses cudaIpcMemHandle_t without including it
#if defined(__has_include)
#if __has_include(<cuda_runtime_api.h>)
#include <cuda_runtime_api.h>
#endif
#endif
#include "../tb200_client.h"
namespace triton { namespace client = ::tb200::client; }
