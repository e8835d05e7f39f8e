/* tb200_grpc_channel.h -- C entry points of libtb200client.so over its gRPC channel (cleartext
 * HTTP/2, csrc/h2.h): a blocking unary call with serialised messages in and out.  Used by the
 * Python drop-in's opt-in native transport (client_b200.grpc.InferenceServerClient(...,
 * transport="native")), where it replaces grpcio's C core on the infer() path
 * (src/python/library/tritonclient/grpc/_client.py:1445-1572). */
#ifndef TB200_CPP_GRPC_CHANNEL_H_
#define TB200_CPP_GRPC_CHANNEL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tb200c_grpc_channel tb200c_grpc_channel;

int tb200c_grpc_channel_open(const char* url, tb200c_grpc_channel** out);
void tb200c_grpc_channel_close(tb200c_grpc_channel* channel);
/* Returns the grpc status (0 = OK).  On success *response is a malloc'ed copy of the response
 * message (release with tb200c_free); otherwise `message` holds grpc-message / the transport
 * error.  `metadata`: 2 * metadata_pairs strings (name, value, ...).  timeout_us 0 = none.
 * Thread safe; calls from several threads share one connection. */
int tb200c_grpc_unary(tb200c_grpc_channel* channel, const char* path, const uint8_t* request, uint64_t request_bytes,
                      const char* const* metadata, int metadata_pairs, uint64_t timeout_us, uint8_t** response,
                      uint64_t* response_bytes, char* message, uint64_t message_cap);
void tb200c_free(void* p);

#ifdef __cplusplus
}
#endif

#endif /* TB200_CPP_GRPC_CHANNEL_H_ */
