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
#include "http_server.h"

namespace tb200 {
void set_last_error(const char* msg);
}

namespace {

// One request of a batch.  kind 0: densenet_onnx (a = data_0, c = fc6_1); kind 1: simple
// (a, b = INPUT0/1, c, d = OUTPUT0/1).
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
    if (j < 16) {
      const int x = static_cast<const int*>(job.a)[j], y = static_cast<const int*>(job.b)[j];
      static_cast<int*>(job.c)[j] = x + y;
      static_cast<int*>(job.d)[j] = x - y;
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
  uint64_t conn_id;
  ModelJob job;
};
constexpr size_t kMaxBatch = 1024;
}  // namespace

struct tb200_mock_server {
  tb200::EpollHttpServer http;
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

void submit(tb200_mock_server* s, uint64_t conn_id, const ModelJob& job) {
  bool wake;
  {
    std::lock_guard<std::mutex> lk(s->qmu);
    wake = s->pending.empty();
    s->pending.push_back(Ticket{conn_id, job});
  }
  if (wake) s->qcv.notify_one();
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
      for (size_t i = 0; i < n; ++i) s->jobs[i] = batch[base + i].job;
      mock_models_kernel<<<dim3(4, static_cast<unsigned>(n)), 256, 0, s->stream>>>(s->jobs);
      const bool ok = cudaStreamSynchronize(s->stream) == cudaSuccess;
      if (!ok) cudaGetLastError();
      for (size_t i = 0; i < n; ++i) {
        const Ticket& t = batch[base + i];
        s->http.CompleteLater(t.conn_id, ok ? kInferOk[t.job.kind] : error_response("model execution failed"));
      }
      s->http.Flush();  // one wake-up per event-loop thread for the whole pass
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
