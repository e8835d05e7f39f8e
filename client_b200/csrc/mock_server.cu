// mock_server.cu -- a small native KServe-v2 HTTP server for loopback load runs.
//
// TOOLING, not part of the client data plane: the reference has no server
// and the Python stand-in (client_b200/testing/mock_server.py) tops out at 2-3 k infer/s,
// which hides every client-side difference.  This one serves the CUDA-shared-memory
// subset natively: it opens the client's IPC handles (so it must run in another process
// than the client), runs the "model" as a CUDA kernel on the mapped regions and answers
// with the JSON the protocol prescribes.  Connections are served by a few epoll threads
// (http_server.h); ONE device thread runs every pending request in one kernel launch and
// hands the answers back in one batch.  Models: densenet_onnx (same arithmetic as the
// Python stand-in: fc6_1[j] = mean of data_0 elements i with i % 1000 == j) and simple
// (OUTPUT0 = INPUT0 + INPUT1, OUTPUT1 = INPUT0 - INPUT1, INT32[1,16]).
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tb200_loadgen.h"
#include "../cpp/grpc_service.pb.h"
#include "grpc_server.h"
#include "http_server.h"

namespace tb200 {
void set_last_error(const char* msg);
}

namespace {

// One request of a batch.  kind 0: densenet_onnx (a = data_0, c = fc6_1); kind 1: simple
// (a, b = INPUT0/1, c, d = OUTPUT0/1); kind 2: bert_large (a = input_ids, b = attention_mask,
// INT64[384] each, c = logits FP32[384]); kind 3: llama3_8b (a = input_ids INT32[`pad`], c = one
// INT64: the sum of the ids, from which the token stream is derived).
struct ModelJob {
  const void* a;
  const void* b;
  void* c;
  void* d;
  int kind;
  int pad;
};

constexpr int kDenseIn = 150528, kDenseOut = 1000, kDenseRows = 151;

// grid (4, njobs): every pending request of the batch in ONE launch
__global__ void mock_models_kernel(const ModelJob* __restrict__ jobs) {
  const ModelJob job = jobs[blockIdx.y];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (job.kind == 1) {
    if (j < 16 && job.a != nullptr) {
      const int x = static_cast<const int*>(job.a)[j], y = static_cast<const int*>(job.b)[j];
      static_cast<int*>(job.c)[j] = x + y;
      static_cast<int*>(job.d)[j] = x - y;
    }
    return;
  }
  if (job.kind == 2) {  // the Python mock's bert_large: (ids % 1000) * mask / 1000 in FP32
    if (j < 384) {
      const long long id = static_cast<const long long*>(job.a)[j], m = static_cast<const long long*>(job.b)[j];
      static_cast<float*>(job.c)[j] = (static_cast<float>(id % 1000) * static_cast<float>(m)) / 1000.0f;
    }
    return;
  }
  if (job.kind == 3) {  // sum of the prompt's token ids (first block of the row only)
    if (blockIdx.x != 0) return;
    __shared__ long long part[8];
    long long acc = 0;
    const int* ids = static_cast<const int*>(job.a);
    for (int i = threadIdx.x; i < job.pad; i += blockDim.x) acc += ids[i];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long total = 0;
      for (int w = 0; w < 8; ++w) total += part[w];
      *static_cast<long long*>(job.c) = total;
    }
    return;
  }
  if (j >= kDenseOut) return;
  const float* in = static_cast<const float*>(job.a);
  float acc = 0.0f;
  // row-by-row like numpy's sum(axis=0, dtype=float32); loads issued 8 rows ahead of the adds
  for (int r0 = 0; r0 < kDenseRows; r0 += 8) {
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = (r0 + k) * kDenseOut + j;
      v[k] = (r0 + k < kDenseRows && i < kDenseIn) ? in[i] : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += v[k];
  }
  static_cast<float*>(job.c)[j] = acc / static_cast<float>(kDenseRows);
}

struct Region {
  void* base = nullptr;
  uint64_t size = 0;
  int device = 0;
};

struct ShmRef {
  std::string region;
  uint64_t size = 0, offset = 0;
};

int b64val(char c) {
  if (c >= 'A' && c <= 'Z') return c - 'A';
  if (c >= 'a' && c <= 'z') return c - 'a' + 26;
  if (c >= '0' && c <= '9') return c - '0' + 52;
  if (c == '+') return 62;
  if (c == '/') return 63;
  return -1;
}
std::vector<uint8_t> b64decode(const std::string& s) {
  std::vector<uint8_t> out;
  uint32_t acc = 0;
  int bits = 0;
  for (char c : s) {
    const int v = b64val(c);
    if (v < 0) continue;
    acc = (acc << 6) | static_cast<uint32_t>(v);
    bits += 6;
    if (bits >= 8) {
      bits -= 8;
      out.push_back(static_cast<uint8_t>((acc >> bits) & 0xFF));
    }
  }
  return out;
}

bool find_string(const std::string& js, const char* key, size_t from, size_t until, std::string* out) {
  const std::string pat = std::string("\"") + key + "\":\"";
  const size_t p = js.find(pat, from);
  if (p == std::string::npos || p >= until) return false;
  const size_t b = p + pat.size();
  const size_t e = js.find('"', b);
  if (e == std::string::npos) return false;
  *out = js.substr(b, e - b);
  return true;
}
bool find_number(const std::string& js, const char* key, size_t from, size_t until, uint64_t* out) {
  const std::string pat = std::string("\"") + key + "\":";
  const size_t p = js.find(pat, from);
  if (p == std::string::npos || p >= until) return false;
  *out = strtoull(js.c_str() + p + pat.size(), nullptr, 10);
  return true;
}

// every tensor object that names a shared memory region, in order, within [from, until)
std::vector<ShmRef> shm_refs(const std::string& js, size_t from, size_t until) {
  std::vector<ShmRef> refs;
  size_t pos = from;
  for (;;) {
    const size_t p = js.find("\"shared_memory_region\":\"", pos);
    if (p == std::string::npos || p >= until) break;
    size_t close = js.find('}', p);
    if (close == std::string::npos || close > until) close = until;
    ShmRef r;
    find_string(js, "shared_memory_region", p, close + 1, &r.region);
    find_number(js, "shared_memory_byte_size", p, close + 1, &r.size);
    find_number(js, "shared_memory_offset", p, close + 1, &r.offset);
    refs.push_back(r);
    pos = close;
  }
  return refs;
}

}  // namespace

namespace {
// an inference request parked with the device thread
struct Ticket {
  uint64_t conn_id;  // HTTP connection, or gRPC call id
  ModelJob job;
  // gRPC only
  bool grpc = false, stream = false, empty_final = false;
  int slab = -1;          // wire-mode tensors live in pinned slab `slab`
  bool wire_out = false;  // the outputs travel in the response (second half of the slab)
  int tokens = 0;         // llama3_8b: responses to send
  std::string id;         // request id, echoed
  // a stream's reply that needs no model run (an error message, the end of the stream): it still
  // travels through the device thread's queue so that the replies of a stream keep request order
  std::shared_ptr<tb200::GrpcReply> canned;
};
constexpr size_t kMaxBatch = 1024;
// wire-mode gRPC requests: inputs are copied into, outputs read from, pinned device-mapped slabs
constexpr size_t kSlabBytes = 32768, kSlabOut = 16384, kSlabs = 4096;
}  // namespace

struct tb200_mock_server {
  tb200::EpollHttpServer http;
  tb200::EpollGrpcServer grpc;
  bool grpc_started = false;
  char* slabs = nullptr;      // pinned, device-mapped (UVA: the same pointer on the device)
  std::mutex slab_mu;
  std::vector<int> slab_free;
  int device = 0;
  std::atomic<bool> device_stop{false};
  std::atomic<uint64_t> requests{0};
  std::atomic<uint64_t> batches{0};
  // model execution: ONE device thread runs every pending request in one launch
  std::thread device_thread;
  std::mutex qmu;
  std::condition_variable qcv;
  std::vector<Ticket> pending;
  ModelJob* jobs = nullptr;  // pinned, device-mapped
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::map<std::string, Region> regions;
};

namespace {

using tb200::EpollHttpServer;

std::string error_response(const std::string& msg) { return EpollHttpServer::Response(400, "{\"error\":\"" + msg + "\"}"); }

const char* kDensenetMeta =
    "{\"name\":\"densenet_onnx\",\"versions\":[\"1\"],\"platform\":\"onnxruntime_onnx\",\"inputs\":[{\"name\":\"data_0\","
    "\"datatype\":\"FP32\",\"shape\":[3,224,224]}],\"outputs\":[{\"name\":\"fc6_1\",\"datatype\":\"FP32\",\"shape\":[1000]}]}";
const char* kSimpleMeta =
    "{\"name\":\"simple\",\"versions\":[\"1\"],\"platform\":\"mock\",\"inputs\":[{\"name\":\"INPUT0\",\"datatype\":\"INT32\","
    "\"shape\":[1,16]},{\"name\":\"INPUT1\",\"datatype\":\"INT32\",\"shape\":[1,16]}],\"outputs\":[{\"name\":\"OUTPUT0\","
    "\"datatype\":\"INT32\",\"shape\":[1,16]},{\"name\":\"OUTPUT1\",\"datatype\":\"INT32\",\"shape\":[1,16]}]}";
const std::string kInferOk[2] = {
    EpollHttpServer::Response(200, "{\"model_name\":\"densenet_onnx\",\"model_version\":\"1\",\"outputs\":[{\"name\":\"fc6_1\",\"datatype\":\"FP32\","
                                   "\"shape\":[1000],\"parameters\":{\"shared_memory_byte_size\":4000}}]}"),
    EpollHttpServer::Response(200, "{\"model_name\":\"simple\",\"model_version\":\"1\",\"outputs\":[{\"name\":\"OUTPUT0\",\"datatype\":\"INT32\","
                                   "\"shape\":[1,16],\"parameters\":{\"shared_memory_byte_size\":64}},{\"name\":\"OUTPUT1\",\"datatype\":"
                                   "\"INT32\",\"shape\":[1,16],\"parameters\":{\"shared_memory_byte_size\":64}}]}")};

bool resolve(tb200_mock_server* s, const ShmRef& r, char** ptr) {
  std::lock_guard<std::mutex> lk(s->mu);
  auto it = s->regions.find(r.region);
  if (it == s->regions.end() || r.offset + r.size > it->second.size) return false;
  *ptr = static_cast<char*>(it->second.base) + r.offset;
  return true;
}

void submit_ticket(tb200_mock_server* s, Ticket&& t) {
  bool wake;
  {
    std::lock_guard<std::mutex> lk(s->qmu);
    wake = s->pending.empty();
    s->pending.push_back(std::move(t));
  }
  if (wake) s->qcv.notify_one();
}
void submit(tb200_mock_server* s, uint64_t conn_id, const ModelJob& job) {
  Ticket t;
  t.conn_id = conn_id;
  t.job = job;
  submit_ticket(s, std::move(t));
}

// ---- gRPC side (inference.GRPCInferenceService) ------------------------------------------------
struct ModelInfo {
  const char* name;
  const char* platform;
  bool decoupled;
  struct Tensor {
    const char* name;
    const char* datatype;
    inference::DataType type;
    std::vector<int64_t> shape;
  };
  std::vector<Tensor> inputs, outputs;
};
const std::vector<ModelInfo>& models() {
  static const std::vector<ModelInfo> kModels = {
      {"densenet_onnx", "onnxruntime_onnx", false, {{"data_0", "FP32", inference::TYPE_FP32, {3, 224, 224}}}, {{"fc6_1", "FP32", inference::TYPE_FP32, {1000}}}},
      {"simple", "mock", false, {{"INPUT0", "INT32", inference::TYPE_INT32, {1, 16}}, {"INPUT1", "INT32", inference::TYPE_INT32, {1, 16}}},
       {{"OUTPUT0", "INT32", inference::TYPE_INT32, {1, 16}}, {"OUTPUT1", "INT32", inference::TYPE_INT32, {1, 16}}}},
      {"bert_large", "mock", false, {{"input_ids", "INT64", inference::TYPE_INT64, {1, 384}}, {"attention_mask", "INT64", inference::TYPE_INT64, {1, 384}}},
       {{"logits", "FP32", inference::TYPE_FP32, {1, 384}}}},
      {"llama3_8b", "mock", true, {{"input_ids", "INT32", inference::TYPE_INT32, {1, -1}}}, {{"token", "INT32", inference::TYPE_INT32, {1, 1}}}},
  };
  return kModels;
}
const ModelInfo* find_model(const std::string& name) {
  for (const ModelInfo& m : models()) {
    if (name == m.name) return &m;
  }
  return nullptr;
}

void grpc_error(tb200::GrpcReply* reply, int status, const std::string& msg) {
  reply->messages.clear();
  reply->finish = true;
  reply->status = status;
  reply->status_message = msg;
}
// inside a stream an error is a message, the stream goes on
void stream_error(tb200::GrpcReply* reply, const std::string& id, const std::string& msg) {
  inference::ModelStreamInferResponse r;
  r.set_error_message(msg);
  r.mutable_infer_response()->set_id(id);
  reply->messages.push_back(r.SerializeAsString());
  reply->finish = false;
}

int take_slab(tb200_mock_server* s) {
  std::lock_guard<std::mutex> lk(s->slab_mu);
  if (s->slab_free.empty()) return -1;
  const int i = s->slab_free.back();
  s->slab_free.pop_back();
  return i;
}
void give_slab(tb200_mock_server* s, int i) {
  if (i < 0) return;
  std::lock_guard<std::mutex> lk(s->slab_mu);
  s->slab_free.push_back(i);
}

bool shm_of(const std::map<std::string, inference::InferParameter>& params, ShmRef* ref) {
  const auto region = params.find("shared_memory_region");
  if (region == params.end()) return false;
  ref->region = region->second.string_param();
  const auto size = params.find("shared_memory_byte_size");
  const auto off = params.find("shared_memory_offset");
  ref->size = size == params.end() ? 0 : static_cast<uint64_t>(size->second.int64_param());
  ref->offset = off == params.end() ? 0 : static_cast<uint64_t>(off->second.int64_param());
  return true;
}

// the responses of one finished ticket: formed on the event-loop thread of its connection
void grpc_build_reply(tb200_mock_server* s, const Ticket& t, bool ok, tb200::GrpcReply* out) {
  tb200::GrpcReply& reply = *out;
  reply.finish = !t.stream;
  const ModelInfo& m = models()[static_cast<size_t>(t.job.kind)];
  if (!ok) {
    if (t.stream) stream_error(&reply, t.id, "model execution failed");
    else grpc_error(&reply, 13, "model execution failed");
    give_slab(s, t.slab);
    return;
  }
  auto base = [&](inference::ModelInferResponse* r) {
    r->set_model_name(m.name);
    r->set_model_version("1");
    r->set_id(t.id);
  };
  auto wrap = [&](const inference::ModelInferResponse& r) {
    if (!t.stream) return r.SerializeAsString();
    std::string inner, out;
    r.AppendTo(&inner);
    tb200::pb::put_bytes(&out, 2, inner);  // ModelStreamInferResponse.infer_response
    return out;
  };
  const char* out_bytes = t.wire_out ? s->slabs + static_cast<size_t>(t.slab) * kSlabBytes + kSlabOut : nullptr;
  if (t.job.kind == 3) {  // llama3_8b: token k = (sum(ids) + k) mod 128256
    long long sum = 0;
    memcpy(&sum, out_bytes, 8);
    const long long first = ((sum % 128256) + 128256) % 128256;
    for (int k = 0; k < t.tokens; ++k) {
      inference::ModelInferResponse r;
      base(&r);
      auto* o = r.add_outputs();
      o->set_name("token");
      o->set_datatype("INT32");
      o->add_shape(1);
      o->add_shape(1);
      const int32_t token = static_cast<int32_t>((first + k) % 128256);
      r.add_raw_output_contents(&token, 4);
      (*r.mutable_parameters())["triton_final_response"].set_bool_param(k == t.tokens - 1 && !t.empty_final);
      reply.messages.push_back(wrap(r));
    }
    if (t.empty_final) {
      inference::ModelInferResponse r;
      base(&r);
      (*r.mutable_parameters())["triton_final_response"].set_bool_param(true);
      reply.messages.push_back(wrap(r));
    }
  } else {
    inference::ModelInferResponse r;
    base(&r);
    size_t off = 0;
    for (const auto& spec : m.outputs) {
      auto* o = r.add_outputs();
      o->set_name(spec.name);
      o->set_datatype(spec.datatype);
      size_t n = strcmp(spec.datatype, "INT64") == 0 ? 8 : 4;
      for (int64_t d : spec.shape) {
        o->add_shape(d);
        n *= static_cast<size_t>(d);
      }
      if (out_bytes != nullptr) {  // wire mode: the tensor travels in the response
        r.add_raw_output_contents(out_bytes + off, n);
        off += n;
      }
    }
    reply.messages.push_back(wrap(r));
  }
  give_slab(s, t.slab);
}

// device thread, after the launch: hand the ticket over, the loop thread serialises
void grpc_complete(tb200_mock_server* s, Ticket&& t, bool ok) {
  tb200::GrpcReply reply;
  const uint64_t call_id = t.conn_id;
  auto ticket = std::make_shared<Ticket>(std::move(t));
  reply.build = [s, ticket, ok](tb200::GrpcReply* out) { grpc_build_reply(s, *ticket, ok, out); };
  s->grpc.CompleteLater(call_id, std::move(reply));
}

// ModelInfer / one message of ModelStreamInfer.  true: *reply is the answer; false: parked.
bool grpc_infer(tb200_mock_server* s, uint64_t call_id, const std::string& message, bool stream, tb200::GrpcReply* reply) {
  inference::ModelInferRequest req;
  if (!req.ParseFromString(message)) {
    grpc_error(reply, 3, "malformed ModelInferRequest");
    return true;
  }
  auto fail = [&](int status, const std::string& msg) {
    if (stream) stream_error(reply, req.id(), msg);
    else grpc_error(reply, status, msg);
    return true;
  };
  s->requests.fetch_add(1, std::memory_order_relaxed);
  Ticket t;
  auto fail_free = [&](int status, const std::string& msg) {  // ... after a slab was taken
    give_slab(s, t.slab);
    t.slab = -1;
    return fail(status, msg);
  };
  const ModelInfo* m = find_model(req.model_name());
  if (m == nullptr) return fail(5, "Request for unknown model: '" + req.model_name() + "' is not found");
  if (m->decoupled && !stream) return fail(3, "ModelInfer RPC doesn't support models with decoupled transaction policy");
  if (req.inputs_size() != static_cast<int>(m->inputs.size())) return fail(3, std::string(m->name) + ": expected " + std::to_string(m->inputs.size()) + " inputs");
  t.conn_id = call_id;
  t.grpc = true;
  t.stream = stream;
  t.id = req.id();
  t.job = ModelJob{nullptr, nullptr, nullptr, nullptr, static_cast<int>(m - models().data()), 0};
  const auto empty_final = req.parameters().find("triton_enable_empty_final_response");
  t.empty_final = m->decoupled && empty_final != req.parameters().end() && empty_final->second.bool_param();
  const auto max_tokens = req.parameters().find("max_tokens");
  t.tokens = max_tokens != req.parameters().end() ? static_cast<int>(max_tokens->second.int64_param()) : 4;
  if (t.tokens < 0 || t.tokens > 4096) return fail(3, "max_tokens out of range");
  // inputs: shared memory by name, or bytes of raw_input_contents copied into a pinned slab
  const void* in_ptr[2] = {nullptr, nullptr};
  int raw_index = 0;
  size_t slab_off = 0;
  char* slab = nullptr;
  for (int i = 0; i < req.inputs_size(); ++i) {
    const auto& tensor = req.inputs(i);
    int slot = -1;
    for (size_t k = 0; k < m->inputs.size(); ++k) {
      if (tensor.name() == m->inputs[k].name) slot = static_cast<int>(k);
    }
    if (slot < 0) return fail_free(3, "unexpected input '" + tensor.name() + "' for model '" + m->name + "'");
    size_t expect = strcmp(m->inputs[static_cast<size_t>(slot)].datatype, "INT64") == 0 ? 8 : 4;
    for (int64_t d : tensor.shape()) expect *= static_cast<size_t>(d < 0 ? 0 : d);
    ShmRef ref;
    if (shm_of(tensor.parameters(), &ref)) {
      char* p = nullptr;
      if (ref.size < expect || !resolve(s, ref, &p)) return fail_free(3, "input '" + tensor.name() + "': shared memory region not registered or too small");
      in_ptr[slot] = p;
      continue;
    }
    if (raw_index >= req.raw_input_contents_size()) return fail_free(3, "input '" + tensor.name() + "' has no data (raw_input_contents expected)");
    const std::string& raw = req.raw_input_contents(raw_index++);
    if (raw.size() != expect || expect == 0) {
      return fail_free(3, "input '" + tensor.name() + "': got " + std::to_string(raw.size()) + " bytes, shape needs " + std::to_string(expect));
    }
    if (slab == nullptr) {
      t.slab = take_slab(s);
      if (t.slab < 0) return fail(8, "the server is out of staging slabs");
      slab = s->slabs + static_cast<size_t>(t.slab) * kSlabBytes;
    }
    if (slab_off + raw.size() > kSlabOut) return fail_free(3, "wire-mode inputs of this server are limited to 16 KiB per request; use shared memory");
    memcpy(slab + slab_off, raw.data(), raw.size());
    in_ptr[slot] = slab + slab_off;
    if (t.job.kind == 3) t.job.pad = static_cast<int>(raw.size() / 4);
    slab_off += (raw.size() + 15) & ~static_cast<size_t>(15);
  }
  t.job.a = in_ptr[0];
  t.job.b = in_ptr[1];
  // outputs: requested in shared memory (by name), else in the slab
  void* out_ptr[2] = {nullptr, nullptr};
  bool any_wire_out = false;
  for (size_t k = 0; k < m->outputs.size(); ++k) {
    ShmRef ref;
    bool in_shm = false;
    for (int i = 0; i < req.outputs_size(); ++i) {
      if (req.outputs(i).name() == m->outputs[k].name && shm_of(req.outputs(i).parameters(), &ref)) in_shm = true;
    }
    if (in_shm) {
      char* p = nullptr;
      if (!resolve(s, ref, &p)) return fail_free(3, std::string("output '") + m->outputs[k].name + "': shared memory region not registered");
      out_ptr[k] = p;
    } else {
      any_wire_out = true;
    }
  }
  if (any_wire_out) {
    if (t.job.kind == 0) return fail_free(3, "densenet_onnx outputs of this server go to shared memory");
    for (size_t k = 0; k < m->outputs.size(); ++k) {
      if (out_ptr[k] != nullptr) return fail_free(3, "outputs must be all in shared memory or all in the response");
    }
    if (t.slab < 0) {
      t.slab = take_slab(s);
      if (t.slab < 0) return fail(8, "the server is out of staging slabs");
      slab = s->slabs + static_cast<size_t>(t.slab) * kSlabBytes;
    }
    out_ptr[0] = slab + kSlabOut;
    if (m->outputs.size() > 1) out_ptr[1] = slab + kSlabOut + 64;  // simple: OUTPUT0 | OUTPUT1, 64 B each
  }
  if (t.job.kind == 0 && (in_ptr[0] == nullptr || slab != nullptr)) return fail_free(3, "densenet_onnx inputs of this server come from shared memory");
  t.job.c = out_ptr[0];
  t.job.d = out_ptr[1];
  t.wire_out = any_wire_out;  // inputs on the wire with outputs in shared memory (or the reverse) are fine
  submit_ticket(s, std::move(t));
  return false;
}

bool handle_grpc(tb200_mock_server* s, uint64_t call_id, const std::string& path, std::string&& message, bool is_message, bool half_close,
                 tb200::GrpcReply* reply) {
  static const std::string kPrefix = "/inference.GRPCInferenceService/";
  if (path.compare(0, kPrefix.size(), kPrefix) != 0) {
    grpc_error(reply, 12, "unknown service");
    return true;
  }
  const std::string rpc = path.substr(kPrefix.size());
  if (rpc == "ModelStreamInfer") {
    if (!is_message) {  // half-close without a message: the stream ends behind what is parked
      Ticket fin;
      fin.conn_id = call_id;
      fin.grpc = true;
      fin.stream = true;
      fin.job.kind = -1;
      submit_ticket(s, std::move(fin));
      return false;
    }
    // the trailers follow when the client half-closes; answers of parked messages come first
    // because replies of one call are queued in order
    const bool now = grpc_infer(s, call_id, message, true, reply);
    if (now) {  // an error message: behind the replies of the messages parked before it
      reply->finish = false;
      Ticket err;
      err.conn_id = call_id;
      err.grpc = true;
      err.stream = true;
      err.job.kind = -1;
      err.canned = std::make_shared<tb200::GrpcReply>(std::move(*reply));
      *reply = tb200::GrpcReply();
      submit_ticket(s, std::move(err));
    }
    if (half_close) {  // a message and the half-close in one frame: the stream ends after its replies
      Ticket fin;
      fin.conn_id = call_id;
      fin.grpc = true;
      fin.stream = true;
      fin.job.kind = -1;
      submit_ticket(s, std::move(fin));
    }
    return false;
  }
  if (!is_message) {
    grpc_error(reply, 13, "unary call without a request message");
    return true;
  }
  auto answer = [&](const tb200::pb::Message& m) {
    reply->messages.push_back(m.SerializeAsString());
    reply->finish = true;
    return true;
  };
  if (rpc == "ModelInfer") return grpc_infer(s, call_id, message, false, reply);
  if (rpc == "ServerLive") {
    inference::ServerLiveResponse r;
    r.set_live(true);
    return answer(r);
  }
  if (rpc == "ServerReady") {
    inference::ServerReadyResponse r;
    r.set_ready(true);
    return answer(r);
  }
  if (rpc == "ModelReady") {
    inference::ModelReadyRequest q;
    q.ParseFromString(message);
    inference::ModelReadyResponse r;
    r.set_ready(find_model(q.name()) != nullptr);
    return answer(r);
  }
  if (rpc == "ServerMetadata") {
    inference::ServerMetadataResponse r;
    r.set_name("triton");
    r.set_version("tb200-native-mock");
    r.add_extensions("cuda_shared_memory");
    r.add_extensions("binary_tensor_data");
    return answer(r);
  }
  if (rpc == "ModelMetadata" || rpc == "ModelConfig") {
    inference::ModelMetadataRequest q;
    q.ParseFromString(message);
    const ModelInfo* m = find_model(q.name());
    if (m == nullptr) {
      grpc_error(reply, 5, "Request for unknown model: '" + q.name() + "' is not found");
      return true;
    }
    if (rpc == "ModelMetadata") {
      inference::ModelMetadataResponse r;
      r.set_name(m->name);
      r.add_versions("1");
      r.set_platform(m->platform);
      for (const auto& t : m->inputs) {
        auto* o = r.add_inputs();
        o->set_name(t.name);
        o->set_datatype(t.datatype);
        for (int64_t d : t.shape) o->add_shape(d);
      }
      for (const auto& t : m->outputs) {
        auto* o = r.add_outputs();
        o->set_name(t.name);
        o->set_datatype(t.datatype);
        for (int64_t d : t.shape) o->add_shape(d);
      }
      return answer(r);
    }
    inference::ModelConfigResponse r;
    auto* c = r.mutable_config();
    c->set_name(m->name);
    c->set_platform(m->platform);
    for (const auto& t : m->inputs) {
      auto* o = c->add_input();
      o->set_name(t.name);
      o->set_data_type(t.type);
      for (int64_t d : t.shape) o->add_dims(d);
    }
    for (const auto& t : m->outputs) {
      auto* o = c->add_output();
      o->set_name(t.name);
      o->set_data_type(t.type);
      for (int64_t d : t.shape) o->add_dims(d);
    }
    if (m->decoupled) c->mutable_model_transaction_policy()->set_decoupled(true);
    return answer(r);
  }
  if (rpc == "RepositoryIndex") {
    inference::RepositoryIndexResponse r;
    for (const ModelInfo& m : models()) {
      auto* e = r.add_models();
      e->set_name(m.name);
      e->set_version("1");
      e->set_state("READY");
    }
    return answer(r);
  }
  if (rpc == "CudaSharedMemoryRegister") {
    inference::CudaSharedMemoryRegisterRequest q;
    if (!q.ParseFromString(message) || q.raw_handle().size() != sizeof(cudaIpcMemHandle_t)) {
      grpc_error(reply, 3, "bad raw_handle");
      return true;
    }
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->regions.count(q.name())) {
      grpc_error(reply, 6, "shared memory region '" + q.name() + "' already in manager");
      return true;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, q.raw_handle().data(), sizeof(h));
    Region r;
    r.size = q.byte_size();
    r.device = static_cast<int>(q.device_id());
    cudaSetDevice(r.device);
    if (cudaIpcOpenMemHandle(&r.base, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      grpc_error(reply, 3, "failed to open CUDA IPC handle");
      return true;
    }
    s->regions[q.name()] = r;
    return answer(inference::CudaSharedMemoryRegisterResponse());
  }
  if (rpc == "CudaSharedMemoryUnregister") {
    inference::CudaSharedMemoryUnregisterRequest q;
    q.ParseFromString(message);
    std::lock_guard<std::mutex> lk(s->mu);
    for (auto it = s->regions.begin(); it != s->regions.end();) {
      if (q.name().empty() || it->first == q.name()) {
        cudaSetDevice(it->second.device);
        cudaIpcCloseMemHandle(it->second.base);
        it = s->regions.erase(it);
      } else {
        ++it;
      }
    }
    return answer(inference::CudaSharedMemoryUnregisterResponse());
  }
  if (rpc == "CudaSharedMemoryStatus") {
    inference::CudaSharedMemoryStatusRequest q;
    q.ParseFromString(message);
    inference::CudaSharedMemoryStatusResponse r;
    std::lock_guard<std::mutex> lk(s->mu);
    for (const auto& kv : s->regions) {
      if (!q.name().empty() && kv.first != q.name()) continue;
      auto& e = (*r.mutable_regions())[kv.first];
      e.set_name(kv.first);
      e.set_device_id(static_cast<uint64_t>(kv.second.device));
      e.set_byte_size(kv.second.size);
    }
    return answer(r);
  }
  if (rpc == "SystemSharedMemoryStatus") return answer(inference::SystemSharedMemoryStatusResponse());
  if (rpc == "SystemSharedMemoryUnregister") return answer(inference::SystemSharedMemoryUnregisterResponse());
  if (rpc == "ModelStatistics") return answer(inference::ModelStatisticsResponse());
  grpc_error(reply, 12, "Method " + rpc + " not implemented by the native stand-in server");
  return true;
}

void device_main(tb200_mock_server* s) {
  cudaSetDevice(s->device);
  std::vector<Ticket> batch;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(s->qmu);
      s->qcv.wait(lk, [&] { return s->device_stop.load() || !s->pending.empty(); });
      if (s->device_stop.load()) return;
      batch.swap(s->pending);
    }
    for (size_t base = 0; base < batch.size(); base += kMaxBatch) {
      const size_t n = std::min(kMaxBatch, batch.size() - base);
      for (size_t i = 0; i < n; ++i) {
        s->jobs[i] = batch[base + i].job;
        if (s->jobs[i].kind < 0) s->jobs[i].kind = 1, s->jobs[i].a = nullptr;  // end-of-stream marker: no work
      }
      bool any_http = false, any_grpc = false;
      // markers ride along as kind-1 rows with null pointers; the kernel skips them
      mock_models_kernel<<<dim3(4, static_cast<unsigned>(n)), 256, 0, s->stream>>>(s->jobs);
      const bool ok = cudaStreamSynchronize(s->stream) == cudaSuccess;
      if (!ok) cudaGetLastError();
      for (size_t i = 0; i < n; ++i) {
        Ticket& t = batch[base + i];
        if (t.grpc) {
          any_grpc = true;
          if (t.job.kind < 0) {  // no model run: a canned stream reply, or the end of the stream
            tb200::GrpcReply end;
            end.finish = true;
            s->grpc.CompleteLater(t.conn_id, t.canned ? std::move(*t.canned) : std::move(end));
          } else {
            grpc_complete(s, std::move(t), ok);
          }
        } else {
          any_http = true;
          s->http.CompleteLater(t.conn_id, ok ? kInferOk[t.job.kind] : error_response("model execution failed"));
        }
      }
      if (any_http) s->http.Flush();  // one wake-up per event-loop thread for the whole pass
      if (any_grpc) s->grpc.Flush();
      s->batches.fetch_add(1, std::memory_order_relaxed);
    }
    batch.clear();
  }
}

// true: *resp holds the answer; false: the request went to the device thread
bool handle(tb200_mock_server* s, uint64_t conn_id, const tb200::HttpRequest& req, std::string* resp) {
  const std::string& path = req.path;
  const std::string& body = req.body;
  if (req.method == "GET") {
    if (path.rfind("/v2/health/", 0) == 0) *resp = EpollHttpServer::Response(200, "");
    else if (path == "/v2/models/densenet_onnx") *resp = EpollHttpServer::Response(200, kDensenetMeta);
    else if (path == "/v2/models/simple") *resp = EpollHttpServer::Response(200, kSimpleMeta);
    else if (path == "/v2/models/densenet_onnx/ready" || path == "/v2/models/simple/ready") *resp = EpollHttpServer::Response(200, "");
    else if (path.rfind("/v2/systemsharedmemory", 0) == 0) *resp = EpollHttpServer::Response(200, "[]");
    else if (path.rfind("/v2/cudasharedmemory", 0) == 0) {
      std::string js = "[";
      std::lock_guard<std::mutex> lk(s->mu);
      for (auto& kv : s->regions) {
        if (js.size() > 1) js += ",";
        js += "{\"name\":\"" + kv.first + "\",\"device_id\":" + std::to_string(kv.second.device) +
              ",\"byte_size\":" + std::to_string(kv.second.size) + "}";
      }
      *resp = EpollHttpServer::Response(200, js + "]");
    } else {
      *resp = error_response("unknown endpoint or model");
    }
    return true;
  }
  const std::string reg = "/v2/cudasharedmemory/region/";
  if (path.rfind(reg, 0) == 0) {
    const size_t slash = path.find('/', reg.size());
    const std::string name = path.substr(reg.size(), slash - reg.size());
    const std::string action = slash == std::string::npos ? "" : path.substr(slash + 1);
    if (action == "register") {
      std::string b64;
      uint64_t dev = 0, size = 0;
      if (!find_string(body, "b64", 0, body.size(), &b64) || !find_number(body, "byte_size", 0, body.size(), &size)) {
        *resp = error_response("malformed register request");
        return true;
      }
      find_number(body, "device_id", 0, body.size(), &dev);
      const std::vector<uint8_t> raw = b64decode(b64);
      if (raw.size() != sizeof(cudaIpcMemHandle_t)) {
        *resp = error_response("bad raw_handle");
        return true;
      }
      std::lock_guard<std::mutex> lk(s->mu);
      if (s->regions.count(name)) {
        *resp = error_response("shared memory region '" + name + "' already in manager");
        return true;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, raw.data(), sizeof(h));
      Region r;
      r.size = size;
      r.device = static_cast<int>(dev);
      cudaSetDevice(r.device);
      if (cudaIpcOpenMemHandle(&r.base, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        *resp = error_response("failed to open CUDA IPC handle");
        return true;
      }
      s->regions[name] = r;
      *resp = EpollHttpServer::Response(200, "");
      return true;
    }
    if (action == "unregister") {
      std::lock_guard<std::mutex> lk(s->mu);
      auto it = s->regions.find(name);
      if (it != s->regions.end()) {
        cudaSetDevice(it->second.device);
        cudaIpcCloseMemHandle(it->second.base);
        s->regions.erase(it);
      }
      *resp = EpollHttpServer::Response(200, "");
      return true;
    }
  }
  if (path.rfind("/v2/systemsharedmemory", 0) == 0 && path.size() > 11 && path.compare(path.size() - 11, 11, "/unregister") == 0) {
    *resp = EpollHttpServer::Response(200, "");  // no system regions are ever registered here
    return true;
  }
  if (path == "/v2/cudasharedmemory/unregister") {
    std::lock_guard<std::mutex> lk(s->mu);
    for (auto& kv : s->regions) {
      cudaSetDevice(kv.second.device);
      cudaIpcCloseMemHandle(kv.second.base);
    }
    s->regions.clear();
    *resp = EpollHttpServer::Response(200, "");
    return true;
  }
  const std::string models = "/v2/models/";
  if (path.rfind(models, 0) == 0 && path.size() > 6 && path.compare(path.size() - 6, 6, "/infer") == 0) {
    const std::string model = path.substr(models.size(), path.size() - 6 - models.size());
    const size_t outs = body.find("\"outputs\":[");
    const std::vector<ShmRef> in = shm_refs(body, 0, outs == std::string::npos ? body.size() : outs);
    const std::vector<ShmRef> out = outs == std::string::npos ? std::vector<ShmRef>() : shm_refs(body, outs, body.size());
    s->requests.fetch_add(1, std::memory_order_relaxed);
    if (model == "densenet_onnx") {
      char *pin, *pout;
      if (in.size() != 1 || out.size() != 1 || in[0].size != 602112 || out[0].size < 4000 || !resolve(s, in[0], &pin) ||
          !resolve(s, out[0], &pout)) {
        *resp = error_response("densenet_onnx: expected data_0 / fc6_1 in registered cuda shared memory");
        return true;
      }
      submit(s, conn_id, ModelJob{pin, nullptr, pout, nullptr, 0, 0});
      return false;
    }
    if (model == "simple") {
      char *a, *b, *c, *d;
      if (in.size() != 2 || out.size() != 2 || in[0].size != 64 || in[1].size != 64 || out[0].size < 64 || out[1].size < 64 ||
          !resolve(s, in[0], &a) || !resolve(s, in[1], &b) || !resolve(s, out[0], &c) || !resolve(s, out[1], &d)) {
        *resp = error_response("simple: expected INPUT0/INPUT1/OUTPUT0/OUTPUT1 in registered cuda shared memory");
        return true;
      }
      submit(s, conn_id, ModelJob{a, b, c, d, 1, 0});
      return false;
    }
    *resp = error_response("Request for unknown model: '" + model + "' is not found");
    return true;
  }
  *resp = error_response("unknown endpoint");
  return true;
}

}  // namespace

extern "C" {

int tb200_mock_server_start(const char* host, int* port, int device_id, tb200_mock_server** out) {
  return tb200_mock_server_start2(host, port, nullptr, device_id, out);
}

int tb200_mock_server_start2(const char* host, int* port, int* grpc_port, int device_id, tb200_mock_server** out) {
  if (host == nullptr || port == nullptr || out == nullptr) {
    tb200::set_last_error("NULL argument");
    return TB200_ERR_INVALID;
  }
  if (cudaSetDevice(device_id) != cudaSuccess) {
    tb200::set_last_error("no usable CUDA device for the mock server");
    return TB200_ERR_CUDA;
  }
  tb200_mock_server* s = new tb200_mock_server();
  s->device = device_id;
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void**>(&s->jobs), kMaxBatch * sizeof(ModelJob), cudaHostAllocMapped) != cudaSuccess) {
    delete s;
    tb200::set_last_error("mock server: CUDA stream / pinned job table allocation failed");
    return TB200_ERR_CUDA;
  }
  s->device_thread = std::thread(device_main, s);
  const int hw = static_cast<int>(std::max(2u, std::thread::hardware_concurrency()));
  if (!s->http.Start(host, port, std::min(16, hw / 2),
                     [s](uint64_t conn_id, const tb200::HttpRequest& req, std::string* resp) { return handle(s, conn_id, req, resp); })) {
    tb200_mock_server_stop(s);
    tb200::set_last_error("cannot bind the mock server");
    return TB200_ERR_IO;
  }
  if (grpc_port != nullptr) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&s->slabs), kSlabs * kSlabBytes, cudaHostAllocMapped) != cudaSuccess) {
      tb200_mock_server_stop(s);
      tb200::set_last_error("mock server: pinned slab allocation failed");
      return TB200_ERR_CUDA;
    }
    for (int i = static_cast<int>(kSlabs) - 1; i >= 0; --i) s->slab_free.push_back(i);
    if (!s->grpc.Start(host, grpc_port, std::min(16, hw / 2),
                       [s](uint64_t call_id, const std::string& path, std::string&& message, bool is_message, bool half_close,
                           tb200::GrpcReply* reply) { return handle_grpc(s, call_id, path, std::move(message), is_message, half_close, reply); })) {
      tb200_mock_server_stop(s);
      tb200::set_last_error("cannot bind the mock server's gRPC port");
      return TB200_ERR_IO;
    }
    s->grpc_started = true;
  }
  *out = s;
  return TB200_OK;
}

uint64_t tb200_mock_server_requests(tb200_mock_server* s) { return s ? s->requests.load() : 0; }
uint64_t tb200_mock_server_batches(tb200_mock_server* s) { return s ? s->batches.load() : 0; }

int tb200_mock_server_stop(tb200_mock_server* s) {
  if (s == nullptr) return TB200_OK;
  {
    std::lock_guard<std::mutex> lk(s->qmu);
    s->device_stop.store(true);
  }
  s->qcv.notify_all();
  if (s->device_thread.joinable()) s->device_thread.join();  // no CompleteLater() after this
  s->http.Stop();                                             // no handler after this
  if (s->grpc_started) s->grpc.Stop();
  if (s->slabs) cudaFreeHost(s->slabs);
  for (auto& kv : s->regions) {
    cudaSetDevice(kv.second.device);
    cudaIpcCloseMemHandle(kv.second.base);
  }
  cudaSetDevice(s->device);
  if (s->jobs) cudaFreeHost(s->jobs);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return TB200_OK;
}

}  // extern "C"
