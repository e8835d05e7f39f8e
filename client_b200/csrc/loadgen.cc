// loadgen.cc -- native closed-loop load generator (C ABI: include/tb200_loadgen.h).
//
// Restates perf_analyzer's ConcurrencyManager / ConcurrencyWorker behaviour (not in the
// reference) with the reference C++ client's RequestTimers / InferStat
// bookkeeping (src/c++/library/common.h:568-648, :93-114; common.cc:56-106).  Transport
// is plain HTTP/1.1 over POSIX sockets: the reference's libcurl path
// (src/c++/library/http_client.cc:1767-1830) cannot be built here and is not needed to
// send a pre-formed request and read a Content-Length response.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tb200_loadgen.h"

namespace {

uint64_t now_ns() {
  return static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(
                                   std::chrono::steady_clock::now().time_since_epoch())
                                   .count());
}

int connect_to(const char* host, int port) {
  int fd = socket(AF_INET, SOCK_STREAM, 0);
  if (fd < 0) return -1;
  sockaddr_in addr{};
  addr.sin_family = AF_INET;
  addr.sin_port = htons(static_cast<uint16_t>(port));
  if (inet_pton(AF_INET, host, &addr.sin_addr) != 1 ||
      connect(fd, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0) {
    close(fd);
    return -1;
  }
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  return fd;
}

bool send_all(int fd, const uint8_t* p, size_t n) {
  while (n > 0) {
    ssize_t k = send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) return false;
    p += k;
    n -= static_cast<size_t>(k);
  }
  return true;
}

// case-insensitive search of "content-length:" in the header block
long content_length(const char* hdr, size_t len) {
  static const char key[] = "content-length:";
  for (size_t i = 0; i + sizeof(key) - 1 <= len; ++i) {
    size_t k = 0;
    while (k < sizeof(key) - 1 && (hdr[i + k] | 0x20) == key[k]) ++k;
    if (k == sizeof(key) - 1 && (i == 0 || hdr[i - 1] == '\n')) return strtol(hdr + i + k, nullptr, 10);
  }
  return 0;
}

// Read one HTTP response; returns the status code (0 on transport error).  *first_byte_ns
// receives the time the first byte arrived (RECV_START).
int read_response(int fd, std::vector<char>& buf, uint64_t* first_byte_ns) {
  buf.clear();
  size_t header_end = 0;
  bool first = true;
  char tmp[4096];
  for (;;) {
    ssize_t k = recv(fd, tmp, sizeof(tmp), 0);
    if (k <= 0) return 0;
    if (first) {
      *first_byte_ns = now_ns();
      first = false;
    }
    buf.insert(buf.end(), tmp, tmp + k);
    if (buf.size() >= 4) {
      const size_t from = buf.size() > static_cast<size_t>(k) + 3 ? buf.size() - k - 3 : 0;
      for (size_t i = from; i + 3 < buf.size(); ++i) {
        if (buf[i] == '\r' && buf[i + 1] == '\n' && buf[i + 2] == '\r' && buf[i + 3] == '\n') {
          header_end = i + 4;
          break;
        }
      }
    }
    if (header_end) break;
    if (buf.size() > (1u << 20)) return 0;
  }
  const long body = content_length(buf.data(), header_end);
  size_t have = buf.size() - header_end;
  while (static_cast<long>(have) < body) {
    ssize_t k = recv(fd, tmp, sizeof(tmp), 0);
    if (k <= 0) return 0;
    have += static_cast<size_t>(k);
  }
  int status = 0;
  if (buf.size() > 12 && memcmp(buf.data(), "HTTP/1.", 7) == 0) status = atoi(buf.data() + 9);
  return status;
}

struct SlotQueue {
  std::mutex mu;
  std::condition_variable cv;
  std::deque<int> q;
  void push(int s) {
    {
      std::lock_guard<std::mutex> lk(mu);
      q.push_back(s);
    }
    cv.notify_one();
  }
  // -1 on timeout
  int pop(int timeout_ms) {
    std::unique_lock<std::mutex> lk(mu);
    if (!cv.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return !q.empty(); })) return -1;
    int s = q.front();
    q.pop_front();
    return s;
  }
  // -1 on timeout
  int pop_ns(uint64_t ns) {
    std::unique_lock<std::mutex> lk(mu);
    if (!cv.wait_for(lk, std::chrono::nanoseconds(ns), [&] { return !q.empty(); })) return -1;
    int s = q.front();
    q.pop_front();
    return s;
  }
  size_t drain(std::vector<int>& out) {
    std::lock_guard<std::mutex> lk(mu);
    const size_t n = q.size();
    out.insert(out.end(), q.begin(), q.end());
    q.clear();
    return n;
  }
};

struct WorkerStats {
  std::mutex mu;
  std::vector<uint64_t> latencies;  // REQUEST_START -> REQUEST_END
  uint64_t completed = 0, failed = 0, total_ns = 0, send_ns = 0, recv_ns = 0;
};

}  // namespace

struct tb200_loadgen {
  std::string host;
  int port = 0;
  int concurrency = 0;
  std::vector<std::vector<uint8_t>> requests;
  std::vector<const uint8_t*> tails;  // borrowed (pinned staging), may be empty
  std::vector<uint64_t> tail_sizes;
  bool passthrough = false;           // no device work per request: workers keep their slot
  tb200_ctx* ctx = nullptr;
  std::vector<tb200_fill_job> fill_jobs;
  int fill_per_slot = 0;
  uint64_t seed = 0;
  bool regenerate = false;
  uint64_t device_window_ns = 0;
  std::vector<tb200_check_job> check_jobs;
  int check_per_slot = 0;
  tb200_check_result* results = nullptr;

  std::atomic<bool> stop{false};
  bool started = false;
  SlotQueue ready, returned;
  std::vector<std::thread> threads;
  std::vector<WorkerStats> stats;
  // device thread counters (guarded by dev_mu)
  std::mutex dev_mu;
  uint64_t device_batches = 0, device_slots = 0, nonfinite = 0, mismatches = 0, epoch = 0;
  std::string error;
  uint64_t window_start_ns = 0;
};

namespace {

void worker_main(tb200_loadgen* lg, int index) {
  WorkerStats& st = lg->stats[index];
  int fd = connect_to(lg->host.c_str(), lg->port);
  std::vector<char> buf;
  buf.reserve(8192);
  while (!lg->stop.load(std::memory_order_relaxed)) {
    const int slot = lg->passthrough ? index : lg->ready.pop(50);
    if (slot < 0) continue;
    const std::vector<uint8_t>& req = lg->requests[slot];
    const uint8_t* tail = lg->tails.empty() ? nullptr : lg->tails[slot];
    const uint64_t tail_size = lg->tails.empty() ? 0 : lg->tail_sizes[slot];
    bool ok = false;
    uint64_t t_start = now_ns(), t_send_end = t_start, t_recv_start = t_start, t_end = t_start;
    for (int attempt = 0; attempt < 2 && !ok; ++attempt) {
      if (fd < 0) fd = connect_to(lg->host.c_str(), lg->port);
      if (fd < 0) break;
      t_start = now_ns();  // REQUEST_START == SEND_START (the request is pre-formed)
      if (!send_all(fd, req.data(), req.size()) || (tail_size != 0 && !send_all(fd, tail, tail_size))) {
        close(fd);
        fd = -1;
        continue;
      }
      t_send_end = now_ns();
      const int status = read_response(fd, buf, &t_recv_start);
      t_end = now_ns();
      if (status == 0) {  // connection dropped: reconnect once
        close(fd);
        fd = -1;
        continue;
      }
      ok = (status == 200);
      break;
    }
    {
      std::lock_guard<std::mutex> lk(st.mu);
      if (ok) {
        st.completed += 1;
        st.total_ns += t_end - t_start;
        st.send_ns += t_send_end - t_start;
        st.recv_ns += t_end - t_recv_start;
        st.latencies.push_back(t_end - t_start);
      } else {
        st.failed += 1;
      }
    }
    if (!lg->passthrough) lg->returned.push(slot);
  }
  if (fd >= 0) close(fd);
}

// validate + regenerate the slots that came back, one launch each, then release them
void device_main(tb200_loadgen* lg) {
  std::vector<int> batch;
  std::vector<tb200_fill_job> fills;
  std::vector<tb200_check_job> checks;
  while (!lg->stop.load(std::memory_order_relaxed)) {
    batch.clear();
    const int first = lg->returned.pop(50);
    if (first < 0) continue;
    batch.push_back(first);
    lg->returned.drain(batch);
    // Accumulation window (cfg.device_window_us): without MPS client and server are two CUDA
    // contexts time-sliced on one GPU and every hand-over costs ~100 us, so a pass should then
    // cover the slots that are about to return anyway.  Ends at once when all are back.
    const uint64_t window_ns = lg->device_window_ns;
    if (window_ns != 0 && lg->ctx != nullptr) {
      const uint64_t t0 = now_ns();
      while (batch.size() < static_cast<size_t>(lg->concurrency)) {
        const uint64_t el = now_ns() - t0;
        if (el >= window_ns) break;
        const int s = lg->returned.pop_ns(window_ns - el);
        if (s < 0) break;
        batch.push_back(s);
        lg->returned.drain(batch);
      }
    }
    if (lg->ctx != nullptr) {
      uint64_t bad = 0, mism = 0;
      const bool do_check = lg->check_per_slot > 0;
      const bool do_fill = lg->regenerate && lg->fill_per_slot > 0;
      checks.clear();
      fills.clear();
      if (do_check) {
        for (int s : batch) {
          for (int k = 0; k < lg->check_per_slot; ++k) checks.push_back(lg->check_jobs[s * lg->check_per_slot + k]);
        }
      }
      if (do_fill) {
        for (int s : batch) {
          for (int k = 0; k < lg->fill_per_slot; ++k) fills.push_back(lg->fill_jobs[s * lg->fill_per_slot + k]);
        }
        lg->epoch += 1ull << 20;  // fresh Philox streams for every generation
      }
      // validation of the returned outputs + generation of the next inputs: parallel branches
      // under one sync (tb200_step_sync).  TB200_LOADGEN_DEVICE_MODE (experiments only):
      // 1 = check, sync, fill, sync; 2 = one stream, one sync
      static const int mode = getenv("TB200_LOADGEN_DEVICE_MODE") ? atoi(getenv("TB200_LOADGEN_DEVICE_MODE")) : 0;
      int rc = TB200_OK;
      if (do_fill && mode == 0) {
        rc = tb200_step_sync(lg->ctx, fills.data(), static_cast<int>(fills.size()), lg->seed, lg->epoch,
                             checks.data(), static_cast<int>(checks.size()), lg->results);
      } else {
        if (do_check) {
          rc = tb200_check_async(lg->ctx, checks.data(), static_cast<int>(checks.size()), lg->results);
          if (rc == TB200_OK && (mode == 1 || !do_fill)) rc = tb200_ctx_sync(lg->ctx);
        }
        if (rc == TB200_OK && do_fill) {
          rc = tb200_fill_async(lg->ctx, fills.data(), static_cast<int>(fills.size()), lg->seed, lg->epoch);
          if (rc == TB200_OK) rc = tb200_ctx_sync(lg->ctx);
        }
      }
      if (rc != TB200_OK) {
        std::lock_guard<std::mutex> lk(lg->dev_mu);
        lg->error = tb200_last_error();
      } else {
        for (size_t k = 0; k < checks.size(); ++k) {
          if (checks[k].kind == TB200_CHECK_TOP1) bad += lg->results[k].mismatches;
          else if (checks[k].kind != TB200_CHECK_SUM) mism += lg->results[k].mismatches;
        }
      }
      std::lock_guard<std::mutex> lk(lg->dev_mu);
      lg->device_batches += 1;
      lg->device_slots += batch.size();
      lg->nonfinite += bad;
      lg->mismatches += mism;
    }
    for (int s : batch) lg->ready.push(s);
  }
}

}  // namespace

namespace tb200 {
void set_last_error(const char* msg);  // runtime.cu: feeds tb200_last_error()
}

namespace {
int lg_fail(int code, const char* msg) {
  tb200::set_last_error(msg);
  return code;
}
}  // namespace

extern "C" {

int tb200_loadgen_create(const tb200_loadgen_config* cfg, tb200_loadgen** out) {
  if (cfg == nullptr || out == nullptr || cfg->host == nullptr || cfg->concurrency <= 0 ||
      cfg->requests == nullptr || cfg->request_sizes == nullptr) {
    return lg_fail(TB200_ERR_INVALID, "bad load generator configuration");
  }
  tb200_loadgen* lg = new tb200_loadgen();
  lg->host = cfg->host;
  lg->port = cfg->port;
  lg->concurrency = cfg->concurrency;
  lg->requests.resize(cfg->concurrency);
  for (int s = 0; s < cfg->concurrency; ++s) {
    lg->requests[s].assign(cfg->requests[s], cfg->requests[s] + cfg->request_sizes[s]);
  }
  lg->ctx = cfg->ctx;
  lg->seed = cfg->seed;
  lg->regenerate = cfg->regenerate != 0;
  lg->device_window_ns = 1000ull * cfg->device_window_us;
  if (cfg->ctx != nullptr && cfg->fill_jobs != nullptr && cfg->fill_jobs_per_slot > 0) {
    lg->fill_per_slot = cfg->fill_jobs_per_slot;
    lg->fill_jobs.assign(cfg->fill_jobs, cfg->fill_jobs + static_cast<size_t>(cfg->concurrency) * cfg->fill_jobs_per_slot);
  }
  if (cfg->ctx != nullptr && cfg->check_jobs != nullptr && cfg->check_jobs_per_slot > 0 && cfg->results != nullptr) {
    lg->check_per_slot = cfg->check_jobs_per_slot;
    lg->check_jobs.assign(cfg->check_jobs, cfg->check_jobs + static_cast<size_t>(cfg->concurrency) * cfg->check_jobs_per_slot);
    lg->results = cfg->results;
  }
  if (cfg->tails != nullptr && cfg->tail_sizes != nullptr) {
    lg->tails.assign(cfg->tails, cfg->tails + cfg->concurrency);
    lg->tail_sizes.assign(cfg->tail_sizes, cfg->tail_sizes + cfg->concurrency);
  }
  lg->passthrough = lg->ctx == nullptr || (lg->check_per_slot == 0 && !(lg->regenerate && lg->fill_per_slot > 0));
  lg->stats = std::vector<WorkerStats>(cfg->concurrency);
  *out = lg;
  return TB200_OK;
}

int tb200_loadgen_start(tb200_loadgen* lg) {
  if (lg == nullptr || lg->started) return lg_fail(TB200_ERR_STATE, "load generator already started");
  // initial generation of every slot: one launch
  if (lg->ctx != nullptr && lg->fill_per_slot > 0) {
    if (tb200_fill_async(lg->ctx, lg->fill_jobs.data(), static_cast<int>(lg->fill_jobs.size()), lg->seed, 0) != 0 ||
        tb200_ctx_sync(lg->ctx) != 0) {
      return lg_fail(TB200_ERR_CUDA, tb200_last_error());
    }
  }
  lg->started = true;
  lg->window_start_ns = now_ns();
  if (!lg->passthrough) {
    for (int s = 0; s < lg->concurrency; ++s) lg->ready.push(s);
    lg->threads.emplace_back(device_main, lg);
  }
  for (int i = 0; i < lg->concurrency; ++i) lg->threads.emplace_back(worker_main, lg, i);
  return TB200_OK;
}

int tb200_loadgen_window(tb200_loadgen* lg, double seconds, tb200_loadgen_stats* out) {
  if (lg == nullptr || out == nullptr || !lg->started) return lg_fail(TB200_ERR_STATE, "load generator not running");
  if (seconds > 0) std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
  memset(out, 0, sizeof(*out));
  std::vector<uint64_t> lat;
  for (WorkerStats& st : lg->stats) {
    std::lock_guard<std::mutex> lk(st.mu);
    out->completed_request_count += st.completed;
    out->failed_request_count += st.failed;
    out->cumulative_total_request_time_ns += st.total_ns;
    out->cumulative_send_time_ns += st.send_ns;
    out->cumulative_receive_time_ns += st.recv_ns;
    lat.insert(lat.end(), st.latencies.begin(), st.latencies.end());
    st.latencies.clear();
    st.completed = st.failed = st.total_ns = st.send_ns = st.recv_ns = 0;
  }
  const uint64_t t = now_ns();
  out->window_seconds = static_cast<double>(t - lg->window_start_ns) * 1e-9;
  lg->window_start_ns = t;
  if (!lat.empty()) {
    std::sort(lat.begin(), lat.end());
    auto pct = [&](double p) { return lat[std::min(lat.size() - 1, static_cast<size_t>(p * (lat.size() - 1) + 0.5))]; };
    out->p50_ns = pct(0.50);
    out->p90_ns = pct(0.90);
    out->p95_ns = pct(0.95);
    out->p99_ns = pct(0.99);
    out->min_ns = lat.front();
    out->max_ns = lat.back();
  }
  {
    std::lock_guard<std::mutex> lk(lg->dev_mu);
    out->device_batches = lg->device_batches;
    out->device_slots = lg->device_slots;
    out->nonfinite_outputs = lg->nonfinite;
    out->check_mismatches = lg->mismatches;
    lg->device_batches = lg->device_slots = lg->nonfinite = lg->mismatches = 0;
    if (!lg->error.empty()) return lg_fail(TB200_ERR_CUDA, lg->error.c_str());
  }
  return TB200_OK;
}

int tb200_loadgen_stop(tb200_loadgen* lg) {
  if (lg == nullptr) return TB200_OK;
  lg->stop.store(true);
  for (std::thread& t : lg->threads) {
    if (t.joinable()) t.join();
  }
  lg->threads.clear();
  lg->started = false;
  return TB200_OK;
}

int tb200_loadgen_destroy(tb200_loadgen* lg) {
  if (lg == nullptr) return TB200_OK;
  tb200_loadgen_stop(lg);
  delete lg;
  return TB200_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// canned-response server (tooling for measuring the generator)
// ---------------------------------------------------------------------------------------
struct tb200_stub_server {
  int listen_fd = -1;
  std::atomic<bool> stop{false};
  std::string response;
  std::thread acceptor;
  std::mutex mu;
  std::vector<std::thread> conns;
  std::vector<int> fds;
};

namespace {

void stub_conn(tb200_stub_server* s, int fd) {
  std::vector<char> buf;
  char tmp[8192];
  for (;;) {
    // read one request: headers, then Content-Length bytes of body
    buf.clear();
    size_t header_end = 0;
    while (!header_end) {
      ssize_t k = recv(fd, tmp, sizeof(tmp), 0);
      if (k <= 0) return;  // the fd is closed by tb200_stub_server_stop
      buf.insert(buf.end(), tmp, tmp + k);
      for (size_t i = 0; i + 3 < buf.size(); ++i) {
        if (buf[i] == '\r' && buf[i + 1] == '\n' && buf[i + 2] == '\r' && buf[i + 3] == '\n') {
          header_end = i + 4;
          break;
        }
      }
    }
    long body = content_length(buf.data(), header_end);
    long have = static_cast<long>(buf.size() - header_end);
    while (have < body) {
      ssize_t k = recv(fd, tmp, sizeof(tmp), 0);
      if (k <= 0) return;
      have += k;
    }
    if (!send_all(fd, reinterpret_cast<const uint8_t*>(s->response.data()), s->response.size())) return;
  }
}

void stub_accept(tb200_stub_server* s) {
  while (!s->stop.load()) {
    int fd = accept(s->listen_fd, nullptr, nullptr);
    if (fd < 0) break;
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    std::lock_guard<std::mutex> lk(s->mu);
    s->fds.push_back(fd);
    s->conns.emplace_back(stub_conn, s, fd);
  }
}

}  // namespace

extern "C" {

int tb200_stub_server_start(const char* host, int* port, const char* response_body, tb200_stub_server** out) {
  if (host == nullptr || port == nullptr || out == nullptr) return lg_fail(TB200_ERR_INVALID, "NULL argument");
  tb200_stub_server* s = new tb200_stub_server();
  const std::string body = response_body ? response_body : "{}";
  s->response = "HTTP/1.1 200 OK\r\nContent-Type: application/json\r\nContent-Length: " + std::to_string(body.size()) +
                "\r\n\r\n" + body;
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
    return lg_fail(TB200_ERR_IO, "cannot bind the stub server");
  }
  socklen_t len = sizeof(addr);
  getsockname(s->listen_fd, reinterpret_cast<sockaddr*>(&addr), &len);
  *port = ntohs(addr.sin_port);
  s->acceptor = std::thread(stub_accept, s);
  *out = s;
  return TB200_OK;
}

int tb200_stub_server_stop(tb200_stub_server* s) {
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
  delete s;
  return TB200_OK;
}

}  // extern "C"
