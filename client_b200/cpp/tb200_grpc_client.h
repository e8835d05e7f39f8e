// tb200_grpc_client.h -- C++ gRPC front end: the reference's InferenceServerGrpcClient surface
// (src/c++/library/grpc_client.h:43-642) over this repository's own HTTP/2 framing
// (client_b200/csrc/h2.h) and message classes (grpc_service.pb.h); no grpc++ / libprotobuf.
//
// Same class and method names, argument meaning and error texts as the reference, so code written
// against `namespace tc = triton::client;` with `#include "grpc_client.h"` compiles against
// compat/grpc_client.h.  Differences by design:
//  * cleartext HTTP/2 with prior knowledge only: use_ssl = true reports an Error;
//  * the tensors of a request go from the AppendRaw scatter list straight into the message
//    (one copy; the reference copies into the protobuf and again at serialisation);
//  * request compression (deflate / gzip message encoding) is done by libtb200's device
//    encoder (tb200_deflate_async) and needs a GPU; responses are requested uncompressed;
//  * KeepAliveOptions: keepalive_time_ms / keepalive_timeout_ms / keepalive_permit_without_calls are
//    honoured with HTTP/2 PING frames (a ping that is not acknowledged in time fails the calls in
//    flight with "keepalive watchdog timeout" and the next call reconnects); generic
//    grpc::ChannelArguments are recorded, the GRPC_ARG_KEEPALIVE_* ones among them are honoured too.
#ifndef TB200_CPP_GRPC_CLIENT_H_
#define TB200_CPP_GRPC_CLIENT_H_

#include <climits>
#include <queue>

#include "grpc_service.pb.h"
#include "tb200_client.h"

// grpc/impl/compression_types.h -- the enumerators the reference's signatures name
enum grpc_compression_algorithm {
  GRPC_COMPRESS_NONE = 0,
  GRPC_COMPRESS_DEFLATE,
  GRPC_COMPRESS_GZIP,
  GRPC_COMPRESS_ALGORITHMS_COUNT
};

// grpc/impl/channel_arg_names.h -- the argument keys the reference's examples pass
#define GRPC_ARG_KEEPALIVE_TIME_MS "grpc.keepalive_time_ms"
#define GRPC_ARG_KEEPALIVE_TIMEOUT_MS "grpc.keepalive_timeout_ms"
#define GRPC_ARG_KEEPALIVE_PERMIT_WITHOUT_CALLS "grpc.keepalive_permit_without_calls"
#define GRPC_ARG_HTTP2_MAX_PINGS_WITHOUT_DATA "grpc.http2.max_pings_without_data"
#define GRPC_ARG_DNS_ENABLE_SRV_QUERIES "grpc.dns_enable_srv_queries"
#define GRPC_ARG_MAX_SEND_MESSAGE_LENGTH "grpc.max_send_message_length"
#define GRPC_ARG_MAX_RECEIVE_MESSAGE_LENGTH "grpc.max_receive_message_length"

namespace grpc {
// stand-in for grpc::ChannelArguments (accepted by Create(), values are recorded only)
class ChannelArguments {
 public:
  void SetInt(const std::string& key, int value) { ints_[key] = value; }
  void SetString(const std::string& key, const std::string& value) { strings_[key] = value; }
  void SetMaxSendMessageSize(int size) { ints_["grpc.max_send_message_length"] = size; }
  void SetMaxReceiveMessageSize(int size) { ints_["grpc.max_receive_message_length"] = size; }
  const std::map<std::string, int>& ints() const { return ints_; }
  const std::map<std::string, std::string>& strings() const { return strings_; }

 private:
  std::map<std::string, int> ints_;
  std::map<std::string, std::string> strings_;
};
}  // namespace grpc

namespace tb200 { namespace client {

// grpc_client.h:43-60
struct SslOptions {
  explicit SslOptions() {}
  std::string root_certificates;
  std::string private_key;
  std::string certificate_chain;
};

// grpc_client.h:63-85
struct KeepAliveOptions {
  explicit KeepAliveOptions()
      : keepalive_time_ms(INT_MAX), keepalive_timeout_ms(20000), keepalive_permit_without_calls(false),
        http2_max_pings_without_data(2) {}
  int keepalive_time_ms;
  int keepalive_timeout_ms;
  bool keepalive_permit_without_calls;
  int http2_max_pings_without_data;
};

namespace detail {
class GrpcChannel;
struct GrpcCall;
}  // namespace detail

// grpc_client.h:100-642
class InferenceServerGrpcClient : public InferenceServerClient {
 public:
  ~InferenceServerGrpcClient();

  static Error Create(std::unique_ptr<InferenceServerGrpcClient>* client, const std::string& server_url,
                      bool verbose = false, bool use_ssl = false, const SslOptions& ssl_options = SslOptions(),
                      const KeepAliveOptions& keepalive_options = KeepAliveOptions(),
                      const bool use_cached_channel = true);
  static Error Create(std::unique_ptr<InferenceServerGrpcClient>* client, const std::string& server_url,
                      const grpc::ChannelArguments& channel_args, bool verbose = false, bool use_ssl = false,
                      const SslOptions& ssl_options = SslOptions(), const bool use_cached_channel = true);

  Error IsServerLive(bool* live, const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error IsServerReady(bool* ready, const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error IsModelReady(bool* ready, const std::string& model_name, const std::string& model_version = "",
                     const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error ServerMetadata(inference::ServerMetadataResponse* server_metadata, const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error ModelMetadata(inference::ModelMetadataResponse* model_metadata, const std::string& model_name,
                      const std::string& model_version = "", const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error ModelConfig(inference::ModelConfigResponse* model_config, const std::string& model_name,
                    const std::string& model_version = "", const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error ModelRepositoryIndex(inference::RepositoryIndexResponse* repository_index, const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error LoadModel(const std::string& model_name, const Headers& headers = Headers(),
                  const std::string& config = std::string(), const std::map<std::string, std::vector<char>>& files = {}, const uint64_t timeout_ms = 0);
  Error UnloadModel(const std::string& model_name, const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error ModelInferenceStatistics(inference::ModelStatisticsResponse* infer_stat, const std::string& model_name = "",
                                 const std::string& model_version = "", const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error UpdateTraceSettings(inference::TraceSettingResponse* response, const std::string& model_name = "",
                            const std::map<std::string, std::vector<std::string>>& settings = {},
                            const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error GetTraceSettings(inference::TraceSettingResponse* settings, const std::string& model_name = "",
                         const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error SystemSharedMemoryStatus(inference::SystemSharedMemoryStatusResponse* status, const std::string& region_name = "",
                                 const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error RegisterSystemSharedMemory(const std::string& name, const std::string& key, const size_t byte_size,
                                   const size_t offset = 0, const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error UnregisterSystemSharedMemory(const std::string& name = "", const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error CudaSharedMemoryStatus(inference::CudaSharedMemoryStatusResponse* status, const std::string& region_name = "",
                               const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  // `cuda_shm_handle`: any 64-byte cudaIpcMemHandle_t-compatible object (see tb200_client.h)
  template <typename IpcHandle>
  Error RegisterCudaSharedMemory(const std::string& name, const IpcHandle& cuda_shm_handle, const size_t device_id,
                                 const size_t byte_size, const Headers& headers = Headers(), const uint64_t timeout_ms = 0) {
    static_assert(sizeof(IpcHandle) == 64, "a CUDA IPC memory handle is 64 bytes");
    return RegisterCudaSharedMemoryRaw(name, reinterpret_cast<const uint8_t*>(&cuda_shm_handle), device_id, byte_size, headers, timeout_ms);
  }
  Error RegisterCudaSharedMemoryRaw(const std::string& name, const uint8_t* handle64, const size_t device_id,
                                    const size_t byte_size, const Headers& headers = Headers(), const uint64_t timeout_ms = 0);
  Error UnregisterCudaSharedMemory(const std::string& name = "", const Headers& headers = Headers(), const uint64_t timeout_ms = 0);

  Error Infer(InferResult** result, const InferOptions& options, const std::vector<InferInput*>& inputs,
              const std::vector<const InferRequestedOutput*>& outputs = std::vector<const InferRequestedOutput*>(),
              const Headers& headers = Headers(), grpc_compression_algorithm compression_algorithm = GRPC_COMPRESS_NONE);
  Error AsyncInfer(OnCompleteFn callback, const InferOptions& options, const std::vector<InferInput*>& inputs,
                   const std::vector<const InferRequestedOutput*>& outputs = std::vector<const InferRequestedOutput*>(),
                   const Headers& headers = Headers(),
                   grpc_compression_algorithm compression_algorithm = GRPC_COMPRESS_NONE);
  Error InferMulti(std::vector<InferResult*>* results, const std::vector<InferOptions>& options,
                   const std::vector<std::vector<InferInput*>>& inputs,
                   const std::vector<std::vector<const InferRequestedOutput*>>& outputs =
                       std::vector<std::vector<const InferRequestedOutput*>>(),
                   const Headers& headers = Headers(),
                   grpc_compression_algorithm compression_algorithm = GRPC_COMPRESS_NONE);
  Error AsyncInferMulti(OnMultiCompleteFn callback, const std::vector<InferOptions>& options,
                        const std::vector<std::vector<InferInput*>>& inputs,
                        const std::vector<std::vector<const InferRequestedOutput*>>& outputs =
                            std::vector<std::vector<const InferRequestedOutput*>>(),
                        const Headers& headers = Headers(),
                        grpc_compression_algorithm compression_algorithm = GRPC_COMPRESS_NONE);

  Error StartStream(OnCompleteFn callback, bool enable_stats = true, uint32_t stream_timeout = 0,
                    const Headers& headers = Headers(),
                    grpc_compression_algorithm compression_algorithm = GRPC_COMPRESS_NONE);
  Error StopStream();
  Error AsyncStreamInfer(const InferOptions& options, const std::vector<InferInput*>& inputs,
                         const std::vector<const InferRequestedOutput*>& outputs =
                             std::vector<const InferRequestedOutput*>());

  size_t GetNumCachedChannels() const;

  // the serialised ModelInferRequest of a call, as Infer() would send it (known-answer tests:
  // tests/golden/wire_golden.json holds the reference Python client's bytes for the same calls)
  static Error SerializeInferRequest(std::string* message, const InferOptions& options,
                                     const std::vector<InferInput*>& inputs,
                                     const std::vector<const InferRequestedOutput*>& outputs =
                                         std::vector<const InferRequestedOutput*>());

 private:
  InferenceServerGrpcClient(const std::string& url, bool verbose, bool use_cached_channel, const KeepAliveOptions& keepalive);
  KeepAliveOptions keepalive_;
  Error Channel(std::shared_ptr<detail::GrpcChannel>* channel);
  Error Unary(const char* method, const tb200::pb::Message& request, tb200::pb::Message* response, const Headers& headers,
              uint64_t timeout_us = 0);
  Error StartInfer(std::shared_ptr<detail::GrpcCall>* call, const InferOptions& options,
                   const std::vector<InferInput*>& inputs, const std::vector<const InferRequestedOutput*>& outputs,
                   const Headers& headers, grpc_compression_algorithm compression_algorithm,
                   std::function<void(detail::GrpcCall*)> on_done, RequestTimers* timer);
  void CallbackWorker();
  void Dispatch(std::function<void()> fn);

  std::string url_;
  bool use_cached_channel_;
  std::mutex channel_mu_;
  std::shared_ptr<detail::GrpcChannel> channel_;

  // user callbacks (AsyncInfer, stream responses) run on this thread, never on the I/O thread
  std::thread worker_;
  std::mutex worker_mu_;
  std::condition_variable worker_cv_;
  std::deque<std::function<void()>> worker_jobs_;
  bool exiting_ = false;
  // asynchronous calls in flight: cancelled and waited for by the destructor (their completion
  // handlers run on the channel's I/O thread and touch this object; the channel may be shared
  // with other clients and outlive this one)
  std::mutex calls_mu_;
  std::condition_variable calls_cv_;
  std::map<detail::GrpcCall*, std::shared_ptr<detail::GrpcCall>> active_calls_;

  // bidirectional stream (one at a time)
  OnCompleteFn stream_callback_;
  std::shared_ptr<detail::GrpcCall> stream_call_;
  bool enable_stream_stats_ = true;
  std::mutex stream_mu_;
  std::condition_variable stream_cv_;
  bool stream_done_ = true;
  std::queue<std::unique_ptr<RequestTimers>> ongoing_stream_request_timers_;
};

}}  // namespace tb200::client

#endif  // TB200_CPP_GRPC_CLIENT_H_
