// mock_server.cu -- a small native KServe-v2 HTTP server for loopback load runs.
//
// TOOLING, not part of the client data plane: the reference has no server
// and the Python stand-in (client_b200/testing/mock_server.py) tops out at 2-3 k infer/s,
// which hides every client-side difference.  This one serves the CUDA-shared-memory
// subset natively: it opens the client's IPC handles (so it must run in another process
// than the client), runs the "model" as a CUDA kernel on the mapped regions and answers
// with the JSON the protocol prescribes.  Models: densenet_onnx (same arithmetic as the
// Python stand-in: fc6_1[j] = mean of data_0 elements i with i % 1000 == j) and simple
// (OUTPUT0 = INPUT0 + INPUT1, OUTPUT1 = INPUT0 - INPUT1, INT32[1,16]).
#include <arpa/inet.h>
#include <cuda_runtime.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

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
// a connection thread's request parked with the device thread
struct Ticket {
  ModelJob job;
  std::mutex m;
  std::condition_variable cv;
  int done = 0;  // 1 ok, -1 failed
};
constexpr size_t kMaxBatch = 1024;
}  // namespace

struct tb200_mock_server {
  int listen_fd = -1;
  int device = 0;
  std::atomic<bool> stop{false};
  std::atomic<bool> device_stop{false};  // set once every connection thread is gone
  std::atomic<uint64_t> requests{0};
  std::atomic<uint64_t> batches{0};
  std::thread acceptor;
  // model execution: ONE device thread runs every pending request in one launch
  std::thread device_thread;
  std::mutex qmu;
  std::condition_variable qcv;
  std::vector<Ticket*> pending;
  ModelJob* jobs = nullptr;  // pinned, device-mapped
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::map<std::string, Region> regions;
  std::vector<std::thread> conns;
  std::vector<int> fds;
};

namespace {

bool send_str(int fd, const std::string& s) {
  const char* p = s.data();
  size_t n = s.size();
  while (n > 0) {
    ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) return false;
    p += k;
    n -= static_cast<size_t>(k);
  }
  return true;
}

std::string http_response(int status, const std::string& body) {
  return "HTTP/1.1 " + std::to_string(status) + (status == 200 ? " OK" : " Bad Request") +
         "\r\nContent-Type: application/json\r\nContent-Length: " + std::to_string(body.size()) + "\r\n\r\n" + body;
}
std::string error_body(const std::string& msg) { return "{\"error\":\"" + msg + "\"}"; }

const char* kDensenetMeta =
    "{\"name\":\"densenet_onnx\",\"versions\":[\"1\"],\"platform\":\"onnxruntime_onnx\",\"inputs\":[{\"name\":\"data_0\","
    "\"datatype\":\"FP32\",\"shape\":[3,224,224]}],\"outputs\":[{\"name\":\"fc6_1\",\"datatype\":\"FP32\",\"shape\":[1000]}]}";
const char* kSimpleMeta =
    "{\"name\":\"simple\",\"versions\":[\"1\"],\"platform\":\"mock\",\"inputs\":[{\"name\":\"INPUT0\",\"datatype\":\"INT32\","
    "\"shape\":[1,16]},{\"name\":\"INPUT1\",\"datatype\":\"INT32\",\"shape\":[1,16]}],\"outputs\":[{\"name\":\"OUTPUT0\","
    "\"datatype\":\"INT32\",\"shape\":[1,16]},{\"name\":\"OUTPUT1\",\"datatype\":\"INT32\",\"shape\":[1,16]}]}";

bool resolve(tb200_mock_server* s, const ShmRef& r, char** ptr) {
  std::lock_guard<std::mutex> lk(s->mu);
  auto it = s->regions.find(r.region);
  if (it == s->regions.end() || r.offset + r.size > it->second.size) return false;
  *ptr = static_cast<char*>(it->second.base) + r.offset;
  return true;
}

bool run_model(tb200_mock_server* s, const ModelJob& job) {
  Ticket t;
  t.job = job;
  {
    std::lock_guard<std::mutex> lk(s->qmu);
    s->pending.push_back(&t);
  }
  s->qcv.notify_one();
  std::unique_lock<std::mutex> lk(t.m);
  t.cv.wait(lk, [&] { return t.done != 0; });
  return t.done > 0;
}

void device_main(tb200_mock_server* s) {
  cudaSetDevice(s->device);
  std::vector<Ticket*> batch;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(s->qmu);
      s->qcv.wait(lk, [&] { return s->device_stop.load() || !s->pending.empty(); });
      if (s->pending.empty()) return;  // stopping
      batch.swap(s->pending);
    }
    for (size_t base = 0; base < batch.size(); base += kMaxBatch) {
      const size_t n = std::min(kMaxBatch, batch.size() - base);
      for (size_t i = 0; i < n; ++i) s->jobs[i] = batch[base + i]->job;
      mock_models_kernel<<<dim3(4, static_cast<unsigned>(n)), 256, 0, s->stream>>>(s->jobs);
      const int ok = (cudaStreamSynchronize(s->stream) == cudaSuccess) ? 1 : -1;
      if (ok < 0) cudaGetLastError();
      for (size_t i = 0; i < n; ++i) {
        Ticket* t = batch[base + i];
        {
          std::lock_guard<std::mutex> lk(t->m);
          t->done = ok;
        }
        t->cv.notify_one();
      }
      s->batches.fetch_add(1, std::memory_order_relaxed);
    }
    batch.clear();
  }
}

std::string handle(tb200_mock_server* s, const std::string& method, const std::string& path,
                   const std::string& body) {
  if (method == "GET") {
    if (path.rfind("/v2/health/", 0) == 0) return http_response(200, "");
    if (path == "/v2/models/densenet_onnx") return http_response(200, kDensenetMeta);
    if (path == "/v2/models/simple") return http_response(200, kSimpleMeta);
    if (path.find("/ready") != std::string::npos) return http_response(200, "");
    if (path.rfind("/v2/cudasharedmemory", 0) == 0) {
      std::string js = "[";
      std::lock_guard<std::mutex> lk(s->mu);
      for (auto& kv : s->regions) {
        if (js.size() > 1) js += ",";
        js += "{\"name\":\"" + kv.first + "\",\"device_id\":" + std::to_string(kv.second.device) +
              ",\"byte_size\":" + std::to_string(kv.second.size) + "}";
      }
      return http_response(200, js + "]");
    }
    return http_response(400, error_body("unknown endpoint"));
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
        return http_response(400, error_body("malformed register request"));
      }
      find_number(body, "device_id", 0, body.size(), &dev);
      const std::vector<uint8_t> raw = b64decode(b64);
      if (raw.size() != sizeof(cudaIpcMemHandle_t)) return http_response(400, error_body("bad raw_handle"));
      std::lock_guard<std::mutex> lk(s->mu);
      if (s->regions.count(name)) return http_response(400, error_body("shared memory region '" + name + "' already in manager"));
      cudaIpcMemHandle_t h;
      memcpy(&h, raw.data(), sizeof(h));
      Region r;
      r.size = size;
      r.device = static_cast<int>(dev);
      cudaSetDevice(r.device);
      if (cudaIpcOpenMemHandle(&r.base, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return http_response(400, error_body("failed to open CUDA IPC handle"));
      }
      s->regions[name] = r;
      return http_response(200, "");
    }
    if (action == "unregister") {
      std::lock_guard<std::mutex> lk(s->mu);
      auto it = s->regions.find(name);
      if (it != s->regions.end()) {
        cudaIpcCloseMemHandle(it->second.base);
        s->regions.erase(it);
      }
      return http_response(200, "");
    }
  }
  if (path == "/v2/cudasharedmemory/unregister") {
    std::lock_guard<std::mutex> lk(s->mu);
    for (auto& kv : s->regions) cudaIpcCloseMemHandle(kv.second.base);
    s->regions.clear();
    return http_response(200, "");
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
        return http_response(400, error_body("densenet_onnx: expected data_0 / fc6_1 in registered cuda shared memory"));
      }
      if (!run_model(s, ModelJob{pin, nullptr, pout, nullptr, 0, 0})) return http_response(400, error_body("model execution failed"));
      return http_response(200, "{\"model_name\":\"densenet_onnx\",\"model_version\":\"1\",\"outputs\":[{\"name\":\"fc6_1\",\"datatype\":\"FP32\","
                                "\"shape\":[1000],\"parameters\":{\"shared_memory_byte_size\":4000}}]}");
    }
    if (model == "simple") {
      char *a, *b, *c, *d;
      if (in.size() != 2 || out.size() != 2 || in[0].size != 64 || in[1].size != 64 || out[0].size < 64 || out[1].size < 64 ||
          !resolve(s, in[0], &a) || !resolve(s, in[1], &b) || !resolve(s, out[0], &c) || !resolve(s, out[1], &d)) {
        return http_response(400, error_body("simple: expected INPUT0/INPUT1/OUTPUT0/OUTPUT1 in registered cuda shared memory"));
      }
      if (!run_model(s, ModelJob{a, b, c, d, 1, 0})) return http_response(400, error_body("model execution failed"));
      return http_response(200, "{\"model_name\":\"simple\",\"model_version\":\"1\",\"outputs\":[{\"name\":\"OUTPUT0\",\"datatype\":\"INT32\","
                                "\"shape\":[1,16],\"parameters\":{\"shared_memory_byte_size\":64}},{\"name\":\"OUTPUT1\",\"datatype\":"
                                "\"INT32\",\"shape\":[1,16],\"parameters\":{\"shared_memory_byte_size\":64}}]}");
    }
    return http_response(400, error_body("Request for unknown model: '" + model + "' is not found"));
  }
  return http_response(400, error_body("unknown endpoint"));
}

void conn_main(tb200_mock_server* s, int fd) {
  cudaSetDevice(s->device);
  std::string buf;
  char tmp[16384];
  for (;;) {
    size_t header_end;
    while ((header_end = buf.find("\r\n\r\n")) == std::string::npos) {
      ssize_t k = recv(fd, tmp, sizeof(tmp), 0);
      if (k <= 0) goto done;
      buf.append(tmp, static_cast<size_t>(k));
    }
    {
      header_end += 4;
      size_t clen = 0;
      for (size_t i = 0; i + 15 <= header_end; ++i) {
        if (strncasecmp(buf.c_str() + i, "content-length:", 15) == 0) {
          clen = strtoull(buf.c_str() + i + 15, nullptr, 10);
          break;
        }
      }
      while (buf.size() < header_end + clen) {
        ssize_t k = recv(fd, tmp, sizeof(tmp), 0);
        if (k <= 0) goto done;
        buf.append(tmp, static_cast<size_t>(k));
      }
      const size_t sp1 = buf.find(' ');
      const size_t sp2 = buf.find(' ', sp1 + 1);
      const std::string method = buf.substr(0, sp1);
      std::string path = buf.substr(sp1 + 1, sp2 - sp1 - 1);
      const size_t q = path.find('?');
      if (q != std::string::npos) path.resize(q);
      const std::string body = buf.substr(header_end, clen);
      const std::string resp = handle(s, method, path, body);
      buf.erase(0, header_end + clen);
      if (!send_str(fd, resp)) goto done;
    }
  }
done:
  return;
}

void accept_main(tb200_mock_server* s) {
  while (!s->stop.load()) {
    int fd = accept(s->listen_fd, nullptr, nullptr);
    if (fd < 0) break;
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    std::lock_guard<std::mutex> lk(s->mu);
    s->fds.push_back(fd);
    s->conns.emplace_back(conn_main, s, fd);
  }
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
  s->listen_fd = socket(AF_INET, SOCK_STREAM, 0);
  int one = 1;
  setsockopt(s->listen_fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in addr{};
  addr.sin_family = AF_INET;
  addr.sin_port = htons(static_cast<uint16_t>(*port));
  if (s->listen_fd < 0 || inet_pton(AF_INET, host, &addr.sin_addr) != 1 ||
      bind(s->listen_fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0 || listen(s->listen_fd, 1024) != 0) {
    if (s->listen_fd >= 0) close(s->listen_fd);
    delete s;
    tb200::set_last_error("cannot bind the mock server");
    return TB200_ERR_IO;
  }
  socklen_t len = sizeof(addr);
  getsockname(s->listen_fd, reinterpret_cast<sockaddr*>(&addr), &len);
  *port = ntohs(addr.sin_port);
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void**>(&s->jobs), kMaxBatch * sizeof(ModelJob), cudaHostAllocMapped) != cudaSuccess) {
    close(s->listen_fd);
    delete s;
    tb200::set_last_error("mock server: CUDA stream / pinned job table allocation failed");
    return TB200_ERR_CUDA;
  }
  s->device_thread = std::thread(device_main, s);
  s->acceptor = std::thread(accept_main, s);
  *out = s;
  return TB200_OK;
}

uint64_t tb200_mock_server_requests(tb200_mock_server* s) { return s ? s->requests.load() : 0; }
uint64_t tb200_mock_server_batches(tb200_mock_server* s) { return s ? s->batches.load() : 0; }

int tb200_mock_server_stop(tb200_mock_server* s) {
  if (s == nullptr) return TB200_OK;
  s->stop.store(true);
  shutdown(s->listen_fd, SHUT_RDWR);
  close(s->listen_fd);
  if (s->acceptor.joinable()) s->acceptor.join();
  {
    std::lock_guard<std::mutex> lk(s->mu);
    for (int fd : s->fds) shutdown(fd, SHUT_RDWR);
  }
  for (std::thread& t : s->conns) {
    if (t.joinable()) t.join();
  }
  for (int fd : s->fds) close(fd);
  {
    std::lock_guard<std::mutex> lk(s->qmu);
    s->device_stop.store(true);
  }
  s->qcv.notify_all();
  if (s->device_thread.joinable()) s->device_thread.join();
  cudaSetDevice(s->device);
  for (auto& kv : s->regions) cudaIpcCloseMemHandle(kv.second.base);
  if (s->jobs) cudaFreeHost(s->jobs);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return TB200_OK;
}

}  // extern "C"
