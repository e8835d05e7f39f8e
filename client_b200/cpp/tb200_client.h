// tb200_client.h -- C++ front end over libtb200: the reference C++ client's HTTP surface
// (src/c++/library/common.h:61-672, http_client.h:105-651) restated for C++ callers, with
// the AppendRaw scatter list feeding the device packer / CUDA-IPC regions.
//
// Same class and method names, argument meaning and error texts as the reference so that
// code written against `namespace tc = triton::client;` compiles against this header
// (compat/http_client.h, compat/common.h alias the namespace).  Differences by design:
//  * transport is a plain keep-alive HTTP/1.1 socket (no libcurl; TLS reports an Error instead of
//    silently doing something else); compressed request bodies are made by libtb200's device
//    encoder (a CUDA device is needed for CompressionType::GZIP|DEFLATE requests), compressed
//    responses are inflated with zlib;
//  * CudaRegion (bottom of this file) puts tensors into server-mapped CUDA-IPC memory
//    through the C ABI of include/tb200.h -- generated, gathered or packed on the device.
#ifndef TB200_CPP_CLIENT_H_
#define TB200_CPP_CLIENT_H_

#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <ostream>
#include <string>
#include <thread>
#include <vector>

struct tb200_ctx;
struct tb200_region;

namespace tb200 { namespace client {

constexpr char kInferHeaderContentLengthHTTPHeader[] = "Inference-Header-Content-Length";
constexpr char kContentLengthHTTPHeader[] = "Content-Length";

// common.h:61-83 -- value type, empty message means success
class Error {
 public:
  explicit Error(const std::string& msg = "") : msg_(msg) {}
  const std::string& Message() const { return msg_; }
  bool IsOk() const { return msg_.empty(); }
  static const Error Success;
  friend std::ostream& operator<<(std::ostream& out, const Error& err);

 private:
  std::string msg_;
};

// common.h:93-114
struct InferStat {
  size_t completed_request_count = 0;
  uint64_t cumulative_total_request_time_ns = 0;
  uint64_t cumulative_send_time_ns = 0;
  uint64_t cumulative_receive_time_ns = 0;
};

// common.h:568-648 -- the six timestamps of one request
class RequestTimers {
 public:
  enum class Kind { REQUEST_START, REQUEST_END, SEND_START, SEND_END, RECV_START, RECV_END, COUNT__ };
  RequestTimers() { Reset(); }
  void Reset() {
    for (uint64_t& t : timestamps_) t = 0;
  }
  uint64_t Timestamp(Kind kind) const { return timestamps_[static_cast<size_t>(kind)]; }
  uint64_t CaptureTimestamp(Kind kind);
  // UINT64_MAX when either end is missing or out of order
  uint64_t Duration(Kind start, Kind end) const;

 private:
  uint64_t timestamps_[static_cast<size_t>(Kind::COUNT__)];
};

// common.h:155-161
struct RequestParameter {
  std::string name;
  std::string value;
  std::string type;  // "string" | "int" | "bool"
};

// common.h:164-232
struct InferOptions {
  explicit InferOptions(const std::string& model_name) : model_name_(model_name) {}
  std::string model_name_;
  std::string model_version_;
  std::string request_id_;
  uint64_t sequence_id_ = 0;
  std::string sequence_id_str_;
  bool sequence_start_ = false;
  bool sequence_end_ = false;
  uint64_t priority_ = 0;
  uint64_t server_timeout_ = 0;
  uint64_t client_timeout_ = 0;  // microseconds, 0 = none
  bool triton_enable_empty_final_response_ = false;
  std::map<std::string, RequestParameter> request_parameters;
};

// common.h:237-395 -- an input tensor as a scatter list of BORROWED buffers
class InferInput {
 public:
  static Error Create(InferInput** infer_input, const std::string& name, const std::vector<int64_t>& dims,
                      const std::string& datatype);
  const std::string& Name() const { return name_; }
  const std::string& Datatype() const { return datatype_; }
  const std::vector<int64_t>& Shape() const { return shape_; }
  Error SetShape(const std::vector<int64_t>& dims);
  Error Reset();
  Error AppendRaw(const std::vector<uint8_t>& input);
  Error AppendRaw(const uint8_t* input, size_t input_byte_size);
  Error SetSharedMemory(const std::string& name, size_t byte_size, size_t offset = 0);
  bool IsSharedMemory() const { return io_type_ == SHARED_MEMORY; }
  Error SharedMemoryInfo(std::string* name, size_t* byte_size, size_t* offset) const;
  Error AppendFromString(const std::vector<std::string>& input);
  Error RawData(const uint8_t** buf, size_t* byte_size);
  Error ByteSize(size_t* byte_size) const;
  bool BinaryData() const { return binary_data_; }
  Error SetBinaryData(const bool binary_data);

  // iteration over the scatter list (common.cc:245-289)
  Error PrepareForRequest();
  Error GetNext(uint8_t* buf, size_t size, size_t* input_bytes, bool* end_of_input);
  Error GetNext(const uint8_t** buf, size_t* input_bytes, bool* end_of_input);

 private:
  InferInput(const std::string& name, const std::vector<int64_t>& dims, const std::string& datatype);
  std::string name_;
  std::vector<int64_t> shape_;
  std::string datatype_;
  size_t byte_size_ = 0;
  size_t bufs_idx_ = 0, buf_pos_ = 0;
  std::vector<const uint8_t*> bufs_;
  std::vector<size_t> buf_byte_sizes_;
  std::deque<std::string> str_bufs_;  // owns AppendFromString serialisations
  enum IOType { NONE, RAW, SHARED_MEMORY };
  IOType io_type_ = NONE;
  std::string shm_name_;
  size_t shm_offset_ = 0;
  bool binary_data_ = true;
};

// common.h:400-483
class InferRequestedOutput {
 public:
  static Error Create(InferRequestedOutput** infer_output, const std::string& name, const size_t class_count = 0,
                      const std::string& datatype = "");
  const std::string& Name() const { return name_; }
  const std::string& Datatype() const { return datatype_; }
  size_t ClassificationCount() const { return class_count_; }
  Error SetSharedMemory(const std::string& region_name, const size_t byte_size, const size_t offset = 0);
  Error UnsetSharedMemory();
  bool IsSharedMemory() const { return io_type_ == SHARED_MEMORY; }
  Error SharedMemoryInfo(std::string* name, size_t* byte_size, size_t* offset) const;
  bool BinaryData() const { return binary_data_; }
  Error SetBinaryData(const bool binary_data);

 private:
  InferRequestedOutput(const std::string& name, const std::string& datatype, const size_t class_count);
  std::string name_;
  std::string datatype_;
  size_t class_count_;
  enum IOType { NONE, SHARED_MEMORY };
  IOType io_type_ = NONE;
  std::string shm_name_;
  size_t shm_byte_size_ = 0;
  size_t shm_offset_ = 0;
  bool binary_data_ = true;
};

// common.h:488-563
class InferResult {
 public:
  virtual ~InferResult() = default;
  virtual Error ModelName(std::string* name) const = 0;
  virtual Error ModelVersion(std::string* version) const = 0;
  virtual Error Id(std::string* id) const = 0;
  virtual Error Shape(const std::string& output_name, std::vector<int64_t>* shape) const = 0;
  virtual Error Datatype(const std::string& output_name, std::string* datatype) const = 0;
  virtual Error RawData(const std::string& output_name, const uint8_t** buf, size_t* byte_size) const = 0;
  virtual Error IsFinalResponse(bool* is_final_response) const = 0;
  virtual Error IsNullResponse(bool* is_null_response) const = 0;
  virtual Error StringData(const std::string& output_name, std::vector<std::string>* string_result) const = 0;
  virtual std::string DebugString() const = 0;
  virtual Error RequestStatus() const = 0;
};

// common.h:119-151
class InferenceServerClient {
 public:
  using OnCompleteFn = std::function<void(InferResult*)>;
  using OnMultiCompleteFn = std::function<void(std::vector<InferResult*>)>;
  explicit InferenceServerClient(bool verbose) : verbose_(verbose) {}
  virtual ~InferenceServerClient() = default;
  Error ClientInferStat(InferStat* infer_stat) const;

 protected:
  Error UpdateInferStat(const RequestTimers& timer);  // common.cc:56-106
  bool verbose_;
  mutable std::mutex stat_mu_;
  InferStat infer_stat_;
};

using Headers = std::map<std::string, std::string>;
using Parameters = std::map<std::string, std::string>;

// http_client.h:45-98 -- accepted for source compatibility; TLS is not built in
struct HttpSslOptions {
  enum CERTTYPE { CERT_PEM = 0, CERT_DER = 1 };
  enum KEYTYPE { KEY_PEM = 0, KEY_DER = 1 };
  long verify_peer = 1;
  long verify_host = 2;
  std::string ca_info;
  CERTTYPE cert_type = CERT_PEM;
  std::string cert;
  KEYTYPE key_type = KEY_PEM;
  std::string key;
};

namespace detail {
class HttpConnection;
// request-side conversions, exposed for the known-answer tests
// (http_client.cc:581-678, cc_client_test.cc:1662-1960)
Error BinaryInputToJsonText(const uint8_t* buf, size_t element_count, const std::string& datatype,
                            std::vector<std::string>* items);
Error BinaryInputsToJsonText(InferInput& input, std::vector<std::string>* items);
// body compression (http_client.cc:146-254): requests are deflated on the device
// (tb200_deflate_async, zlib / gzip container), responses inflated with zlib on the host
Error DeflateOnDevice(std::string* body, bool gzip);
Error Inflate(const std::string& in, std::string* out);
}  // namespace detail

// http_client.h:105-651
class InferenceServerHttpClient : public InferenceServerClient {
 public:
  enum class CompressionType { NONE, DEFLATE, GZIP };
  ~InferenceServerHttpClient();

  static Error GenerateRequestBody(std::vector<char>* request_body, size_t* header_length, const InferOptions& options,
                                   const std::vector<InferInput*>& inputs,
                                   const std::vector<const InferRequestedOutput*>& outputs =
                                       std::vector<const InferRequestedOutput*>());
  static Error ParseResponseBody(InferResult** result, const std::vector<char>& response_body,
                                 size_t header_length = 0);

  static Error Create(std::unique_ptr<InferenceServerHttpClient>* client, const std::string& server_url,
                      bool verbose = false, const HttpSslOptions& ssl_options = HttpSslOptions());

  Error IsServerLive(bool* live, const Headers& headers = Headers(), const Parameters& query_params = Parameters());
  Error IsServerReady(bool* ready, const Headers& headers = Headers(), const Parameters& query_params = Parameters());
  Error IsModelReady(bool* ready, const std::string& model_name, const std::string& model_version = "",
                     const Headers& headers = Headers(), const Parameters& query_params = Parameters());
  Error ServerMetadata(std::string* server_metadata, const Headers& headers = Headers(),
                       const Parameters& query_params = Parameters());
  Error ModelMetadata(std::string* model_metadata, const std::string& model_name,
                      const std::string& model_version = "", const Headers& headers = Headers(),
                      const Parameters& query_params = Parameters());
  Error ModelConfig(std::string* model_config, const std::string& model_name, const std::string& model_version = "",
                    const Headers& headers = Headers(), const Parameters& query_params = Parameters());
  Error ModelRepositoryIndex(std::string* repository_index, const Headers& headers = Headers(),
                             const Parameters& query_params = Parameters());
  Error LoadModel(const std::string& model_name, const Headers& headers = Headers(),
                  const Parameters& query_params = Parameters(), const std::string& config = std::string(),
                  const std::map<std::string, std::vector<char>>& files = {});
  Error UnloadModel(const std::string& model_name, const Headers& headers = Headers(),
                    const Parameters& query_params = Parameters());
  Error ModelInferenceStatistics(std::string* infer_stat, const std::string& model_name = "",
                                 const std::string& model_version = "", const Headers& headers = Headers(),
                                 const Parameters& query_params = Parameters());
  Error UpdateTraceSettings(std::string* response, const std::string& model_name = "",
                            const std::map<std::string, std::vector<std::string>>& settings = {},
                            const Headers& headers = Headers(), const Parameters& query_params = Parameters());
  Error GetTraceSettings(std::string* settings, const std::string& model_name = "", const Headers& headers = Headers(),
                         const Parameters& query_params = Parameters());
  Error SystemSharedMemoryStatus(std::string* status, const std::string& region_name = "",
                                 const Headers& headers = Headers(), const Parameters& query_params = Parameters());
  Error RegisterSystemSharedMemory(const std::string& name, const std::string& key, const size_t byte_size,
                                   const size_t offset = 0, const Headers& headers = Headers(),
                                   const Parameters& query_params = Parameters());
  Error UnregisterSystemSharedMemory(const std::string& name = "", const Headers& headers = Headers(),
                                     const Parameters& query_params = Parameters());
  Error CudaSharedMemoryStatus(std::string* status, const std::string& region_name = "",
                               const Headers& headers = Headers(), const Parameters& query_params = Parameters());
  // `cuda_shm_handle`: any 64-byte cudaIpcMemHandle_t-compatible object (the reference
  // takes `const cudaIpcMemHandle_t&`; a template keeps this header free of CUDA includes)
  template <typename IpcHandle>
  Error RegisterCudaSharedMemory(const std::string& name, const IpcHandle& cuda_shm_handle, const size_t device_id,
                                 const size_t byte_size, const Headers& headers = Headers(),
                                 const Parameters& query_params = Parameters()) {
    static_assert(sizeof(IpcHandle) == 64, "a CUDA IPC memory handle is 64 bytes");
    return RegisterCudaSharedMemoryRaw(name, reinterpret_cast<const uint8_t*>(&cuda_shm_handle), device_id, byte_size,
                                       headers, query_params);
  }
  Error RegisterCudaSharedMemoryRaw(const std::string& name, const uint8_t* handle64, const size_t device_id,
                                    const size_t byte_size, const Headers& headers = Headers(),
                                    const Parameters& query_params = Parameters());
  Error UnregisterCudaSharedMemory(const std::string& name = "", const Headers& headers = Headers(),
                                   const Parameters& query_params = Parameters());

  Error Infer(InferResult** result, const InferOptions& options, const std::vector<InferInput*>& inputs,
              const std::vector<const InferRequestedOutput*>& outputs = std::vector<const InferRequestedOutput*>(),
              const Headers& headers = Headers(), const Parameters& query_params = Parameters(),
              const CompressionType request_compression_algorithm = CompressionType::NONE,
              const CompressionType response_compression_algorithm = CompressionType::NONE);
  Error AsyncInfer(OnCompleteFn callback, const InferOptions& options, const std::vector<InferInput*>& inputs,
                   const std::vector<const InferRequestedOutput*>& outputs = std::vector<const InferRequestedOutput*>(),
                   const Headers& headers = Headers(), const Parameters& query_params = Parameters(),
                   const CompressionType request_compression_algorithm = CompressionType::NONE,
                   const CompressionType response_compression_algorithm = CompressionType::NONE);
  Error InferMulti(std::vector<InferResult*>* results, const std::vector<InferOptions>& options,
                   const std::vector<std::vector<InferInput*>>& inputs,
                   const std::vector<std::vector<const InferRequestedOutput*>>& outputs =
                       std::vector<std::vector<const InferRequestedOutput*>>(),
                   const Headers& headers = Headers(), const Parameters& query_params = Parameters(),
                   const CompressionType request_compression_algorithm = CompressionType::NONE,
                   const CompressionType response_compression_algorithm = CompressionType::NONE);
  Error AsyncInferMulti(OnMultiCompleteFn callback, const std::vector<InferOptions>& options,
                        const std::vector<std::vector<InferInput*>>& inputs,
                        const std::vector<std::vector<const InferRequestedOutput*>>& outputs =
                            std::vector<std::vector<const InferRequestedOutput*>>(),
                        const Headers& headers = Headers(), const Parameters& query_params = Parameters(),
                        const CompressionType request_compression_algorithm = CompressionType::NONE,
                        const CompressionType response_compression_algorithm = CompressionType::NONE);

 private:
  InferenceServerHttpClient(const std::string& host, int port, const std::string& base_path, bool verbose);
  struct AsyncJob;
  Error Get(const std::string& path, const Headers& headers, const Parameters& query_params, std::string* response,
            long* http_code = nullptr);
  Error Post(const std::string& path, const std::string& request, const Headers& headers,
             const Parameters& query_params, std::string* response, long* http_code = nullptr);
  Error InferOn(detail::HttpConnection* conn, InferResult** result, const InferOptions& options,
                const std::vector<InferInput*>& inputs, const std::vector<const InferRequestedOutput*>& outputs,
                const Headers& headers, const Parameters& query_params, CompressionType request_compression,
                CompressionType response_compression);
  void AsyncWorker();

  std::string host_;
  int port_;
  std::string base_path_;
  std::mutex sync_mu_;  // the reference's easy handle is single-threaded; calls serialise here
  std::unique_ptr<detail::HttpConnection> sync_conn_;
  // asynchronous requests: one worker thread with its own connection (http_client.cc AsyncTransfer)
  std::thread worker_;
  std::mutex async_mu_;
  std::condition_variable async_cv_;
  std::deque<std::shared_ptr<AsyncJob>> async_jobs_;
  bool exiting_ = false;
};

// ---------------------------------------------------------------------------------------
// Device side: a CUDA-IPC region owned by this process, filled through libtb200's C ABI.
// The counterpart of the reference's examples that cudaMalloc + cudaIpcGetMemHandle +
// cudaMemcpy by hand (src/c++/examples/simple_http_cudashm_client.cc:45-75,165-215).
// ---------------------------------------------------------------------------------------
class CudaRegion {
 public:
  ~CudaRegion();
  static Error Create(std::unique_ptr<CudaRegion>* region, const std::string& name, size_t byte_size, int device_id = 0);
  const std::string& Name() const { return name_; }
  size_t ByteSize() const { return byte_size_; }
  int DeviceId() const { return device_id_; }
  void* DevicePtr() const;
  // 64-byte cudaIpcMemHandle_t for RegisterCudaSharedMemoryRaw()
  Error IpcHandle(uint8_t out[64]) const;
  Error Register(InferenceServerHttpClient* client) const;

  // the input's AppendRaw scatter list -> [offset, offset + ByteSize) of the region, one
  // gathered host->device transfer (no intermediate concatenation)
  Error SetFromInput(InferInput& input, size_t offset = 0);
  Error Write(size_t offset, const void* src, size_t byte_size);
  Error Read(size_t offset, void* dst, size_t byte_size) const;
  // synthetic tensor generated in place on the device (fill contract of DESIGN.md section 3):
  // uniform [0,1) for floating types, full range for integers; `zero` for zero data
  Error FillRandom(size_t offset, const std::string& datatype, size_t byte_size, uint64_t seed, uint64_t stream_id,
                   bool zero = false);
  // on-device validation of the add/sub model: OUTPUT0 == IN0 + IN1, OUTPUT1 == IN0 - IN1 (INT32)
  Error CheckAddSub(size_t out0_offset, size_t out1_offset, const CudaRegion& inputs, size_t in0_offset,
                    size_t in1_offset, size_t byte_size, uint64_t* mismatches) const;

 private:
  CudaRegion() = default;
  std::string name_;
  size_t byte_size_ = 0;
  int device_id_ = 0;
  tb200_ctx* ctx_ = nullptr;
  tb200_region* region_ = nullptr;
};

}}  // namespace tb200::client

#endif  // TB200_CPP_CLIENT_H_
