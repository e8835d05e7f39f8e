// Source-compatibility shim for `#include "common.h"` (Error, InferInput, InferOptions, ...).
#pragma once
#include "../tb200_client.h"
namespace triton { namespace client = ::tb200::client; }
