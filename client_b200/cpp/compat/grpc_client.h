// Source-compatibility shim: code written against the reference's `#include "grpc_client.h"` /
// `namespace tc = triton::client;` builds against tb200_grpc_client.h (message classes in
// namespace `inference`, grpc_compression_algorithm and grpc::ChannelArguments stand-ins).
#pragma once
#include "../tb200_grpc_client.h"
namespace triton { namespace client = ::tb200::client; }
