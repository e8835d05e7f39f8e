// Source-compatibility shim: code written against the reference's
// `#include "http_client.h"` / `namespace tc = triton::client;` builds against tb200_client.h.
#pragma once
#include "../tb200_client.h"
namespace triton { namespace client = ::tb200::client; }
