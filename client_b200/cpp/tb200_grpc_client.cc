// tb200_grpc_client.cc -- see tb200_grpc_client.h.  Layout of this file:
//   detail::GrpcChannel   one HTTP/2 connection + its I/O thread (all streams multiplexed)
//   InferResultGrpc       grpc_client.cc:178-446 restated
//   InferenceServerGrpcClient
#include "tb200_grpc_client.h"

#include <arpa/inet.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <iostream>

#include "../csrc/h2.h"
#include "tb200.h"

namespace tb200 { namespace client {

namespace {
using Clock = std::chrono::steady_clock;
constexpr uint32_t kStreamWindow = 4u << 20;   // what we let a peer send per stream before we acknowledge
constexpr uint32_t kConnWindow = 1u << 30;
// grpc status codes this file names (grpc/status.h)
enum GrpcStatus { kOk = 0, kCancelled = 1, kUnknown = 2, kDeadlineExceeded = 4, kResourceExhausted = 8, kUnimplemented = 12, kInternal = 13, kUnavailable = 14 };

std::string PercentDecode(const std::string& s) {  // grpc-message is percent-encoded
  std::string out;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '%' && i + 2 < s.size() + 0 && isxdigit(static_cast<unsigned char>(s[i + 1])) && isxdigit(static_cast<unsigned char>(s[i + 2]))) {
      out.push_back(static_cast<char>(std::stoi(s.substr(i + 1, 2), nullptr, 16)));
      i += 2;
    } else {
      out.push_back(s[i]);
    }
  }
  return out;
}
std::string Lower(std::string s) {
  for (char& c : s) c = static_cast<char>(tolower(static_cast<unsigned char>(c)));
  return s;
}
}  // namespace

namespace detail {

struct GrpcCall {
  // ---- set before Start
  std::string path;
  Headers metadata;
  uint64_t timeout_us = 0;
  std::string encoding;  // "" | "deflate" | "gzip": grpc-encoding of the messages we send
  // every complete message of the response, on the I/O thread
  std::function<void(std::string&&)> on_message;
  // once, on the I/O thread, after the last on_message
  std::function<void(GrpcCall*)> on_done;
  // ---- result
  int status = -1;
  std::string status_message;
  std::string response;  // unary calls without on_message: the (last) response message
  // ---- I/O thread state
  uint32_t stream_id = 0;
  std::deque<std::string> pending;  // framed messages (5-byte prefix + body) not fully sent
  size_t pending_off = 0;
  bool writes_done = false, end_sent = false, finished = false;
  int64_t send_window = 0;
  uint32_t recv_consumed = 0;
  std::string rx;
  Clock::time_point deadline{};
  bool has_deadline = false;
  // ---- for blocking callers
  std::mutex mu;
  std::condition_variable cv;
  bool done = false;
  void Wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [this] { return done; });
  }
};

struct ChannelKeepAlive {
  int time_ms = INT_MAX;       // silence from the peer after which a PING goes out
  int timeout_ms = 20000;      // how long its acknowledgement may take
  bool without_calls = false;  // ping even when no call is in flight
};

class GrpcChannel {
 public:
  using KeepAlive = ChannelKeepAlive;
  static Error Connect(const std::string& url, std::shared_ptr<GrpcChannel>* out, const KeepAlive& keepalive = KeepAlive()) {
    std::string host = url;
    std::string port = "80";
    const size_t scheme = host.find("://");
    if (scheme != std::string::npos) host = host.substr(scheme + 3);
    const size_t colon = host.rfind(':');
    if (colon != std::string::npos) {
      port = host.substr(colon + 1);
      host = host.substr(0, colon);
    }
    addrinfo hints{};
    hints.ai_family = AF_UNSPEC;
    hints.ai_socktype = SOCK_STREAM;
    addrinfo* res = nullptr;
    if (getaddrinfo(host.c_str(), port.c_str(), &hints, &res) != 0 || res == nullptr) {
      return Error("failed to connect to all addresses; DNS resolution failed for " + url);
    }
    int fd = -1;
    for (addrinfo* a = res; a != nullptr; a = a->ai_next) {
      fd = socket(a->ai_family, a->ai_socktype, a->ai_protocol);
      if (fd < 0) continue;
      if (connect(fd, a->ai_addr, a->ai_addrlen) == 0) break;
      close(fd);
      fd = -1;
    }
    freeaddrinfo(res);
    if (fd < 0) return Error("failed to connect to all addresses; last error: connection to " + url + " refused");
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK);
    std::shared_ptr<GrpcChannel> ch(new GrpcChannel(fd, host + ":" + port));
    ch->keepalive_ = keepalive;
    ch->io_ = std::thread(&GrpcChannel::IoMain, ch.get());
    *out = std::move(ch);
    return Error::Success;
  }

  ~GrpcChannel() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    Wake();
    if (io_.joinable()) io_.join();
    close(fd_);
    close(evfd_);
  }

  bool Broken() const { return broken_.load(); }

  // start a call; `first_message` (may be empty = none yet) is the unframed request message
  void Start(const std::shared_ptr<GrpcCall>& call, std::string&& framed_first, bool writes_done) {
    if (!framed_first.empty()) call->pending.push_back(std::move(framed_first));
    call->writes_done = writes_done;
    Push(Command{Command::START, call, std::string()});
  }
  void Write(const std::shared_ptr<GrpcCall>& call, std::string&& framed) { Push(Command{Command::WRITE, call, std::move(framed)}); }
  void WritesDone(const std::shared_ptr<GrpcCall>& call) { Push(Command{Command::WRITES_DONE, call, std::string()}); }
  void Cancel(const std::shared_ptr<GrpcCall>& call) { Push(Command{Command::CANCEL, call, std::string()}); }

  // 5-byte gRPC prefix + body
  static std::string Framed(const std::string& body, bool compressed = false) {
    std::string m(5, '\0');
    h2::put_grpc_prefix(reinterpret_cast<uint8_t*>(&m[0]), static_cast<uint32_t>(body.size()));
    m[0] = compressed ? 1 : 0;
    m += body;
    return m;
  }

 private:
  struct Command {
    enum Kind { START, WRITE, WRITES_DONE, CANCEL } kind;
    std::shared_ptr<GrpcCall> call;
    std::string data;
  };

  GrpcChannel(int fd, const std::string& authority) : fd_(fd), authority_(authority) {
    evfd_ = eventfd(0, EFD_NONBLOCK);
    static_assert(kConnWindow == h2::kOurConnWindow, "client_preface() opens the connection window");
    out_ = h2::client_preface();  // 1 MiB stream windows, raised here
    out_ += h2::frame(h2::SETTINGS, 0, 0, h2::setting(h2::kSettingsInitialWindow, kStreamWindow));
  }

  void Push(Command&& c) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      commands_.push_back(std::move(c));
    }
    Wake();
  }
  void Wake() {
    const uint64_t one = 1;
    if (write(evfd_, &one, sizeof(one)) < 0) return;
  }

  void Finish(const std::shared_ptr<GrpcCall>& call, int status, const std::string& message) {
    if (call->finished) return;
    call->finished = true;
    call->status = status;
    call->status_message = message;
    if (call->stream_id != 0) streams_.erase(call->stream_id);
    if (call->on_done) call->on_done(call.get());
    {
      std::lock_guard<std::mutex> lk(call->mu);
      call->done = true;
    }
    call->cv.notify_all();
  }

  void FailAll(const std::string& why) {
    broken_.store(true);
    std::vector<std::shared_ptr<GrpcCall>> calls;
    for (auto& kv : streams_) calls.push_back(kv.second);
    for (auto& c : calls) Finish(c, kUnavailable, why);
  }

  void SendHeaders(const std::shared_ptr<GrpcCall>& call) {
    std::string block = h2::grpc_request_headers(authority_, call->path);
    h2::hpack_literal(&block, "grpc-accept-encoding", "identity,deflate,gzip");
    if (!call->encoding.empty()) h2::hpack_literal(&block, "grpc-encoding", call->encoding);
    if (call->timeout_us != 0) h2::hpack_literal(&block, "grpc-timeout", std::to_string(call->timeout_us) + "u");
    for (const auto& kv : call->metadata) h2::hpack_literal(&block, Lower(kv.first), kv.second);
    // header blocks larger than a frame continue in CONTINUATION frames
    size_t off = 0;
    bool first = true;
    do {
      const size_t n = std::min<size_t>(block.size() - off, peer_max_frame_);
      const bool last = off + n == block.size();
      out_ += h2::frame(first ? h2::HEADERS : h2::CONTINUATION, last ? h2::kEndHeaders : 0, call->stream_id, block.substr(off, n));
      off += n;
      first = false;
    } while (off < block.size());
  }

  // as much of the call's pending messages as the windows allow
  void Pump(const std::shared_ptr<GrpcCall>& call) {
    if (call->finished || call->end_sent || call->stream_id == 0) return;
    while (!call->pending.empty()) {
      const std::string& m = call->pending.front();
      const size_t left = m.size() - call->pending_off;
      int64_t room = std::min<int64_t>(call->send_window, conn_send_window_);
      if (room <= 0) return;
      const size_t n = std::min<size_t>(std::min<size_t>(left, static_cast<size_t>(room)), peer_max_frame_);
      const bool last_of_message = n == left;
      const bool end = last_of_message && call->pending.size() == 1 && call->writes_done;
      uint8_t hdr[9];
      h2::put_frame_header(hdr, static_cast<uint32_t>(n), h2::DATA, end ? h2::kEndStream : 0, call->stream_id);
      out_.append(reinterpret_cast<const char*>(hdr), 9);
      out_.append(m, call->pending_off, n);
      call->send_window -= static_cast<int64_t>(n);
      conn_send_window_ -= static_cast<int64_t>(n);
      if (last_of_message) {
        call->pending.pop_front();
        call->pending_off = 0;
      } else {
        call->pending_off += n;
      }
      if (end) {
        call->end_sent = true;
        return;
      }
    }
    if (call->writes_done && !call->end_sent) {
      out_ += h2::frame(h2::DATA, h2::kEndStream, call->stream_id, "");
      call->end_sent = true;
    }
  }

  void HandleCommands() {
    std::deque<Command> cmds;
    {
      std::lock_guard<std::mutex> lk(mu_);
      cmds.swap(commands_);
    }
    for (Command& c : cmds) {
      const std::shared_ptr<GrpcCall>& call = c.call;
      switch (c.kind) {
        case Command::START:
          if (broken_.load()) {
            Finish(call, kUnavailable, "Socket closed");
            break;
          }
          call->stream_id = next_stream_id_;
          next_stream_id_ += 2;
          call->send_window = peer_initial_window_;
          if (call->timeout_us != 0) {
            call->has_deadline = true;
            call->deadline = Clock::now() + std::chrono::microseconds(call->timeout_us);
          }
          streams_[call->stream_id] = call;
          SendHeaders(call);
          Pump(call);
          break;
        case Command::WRITE:
          if (call->finished || call->writes_done) break;
          call->pending.push_back(std::move(c.data));
          Pump(call);
          break;
        case Command::WRITES_DONE:
          call->writes_done = true;
          Pump(call);
          break;
        case Command::CANCEL:
          if (!call->finished && call->stream_id != 0) {
            std::string code;
            h2::put_u32(&code, 8);  // CANCEL
            out_ += h2::frame(h2::RST_STREAM, 0, call->stream_id, code);
            Finish(call, kCancelled, "Cancelled");
          }
          break;
      }
    }
  }

  void HandleHeaders(uint32_t stream, const std::string& block, bool end_stream) {
    std::vector<h2::HpackDecoder::Field> fields;
    if (!hpack_.Decode(reinterpret_cast<const uint8_t*>(block.data()), block.size(), &fields)) {
      FailAll("malformed response headers");
      return;
    }
    auto it = streams_.find(stream);
    if (it == streams_.end()) return;
    std::shared_ptr<GrpcCall> call = it->second;
    int grpc_status = -1;
    std::string message, http_status;
    for (const auto& f : fields) {
      if (f.first == "grpc-status") grpc_status = atoi(f.second.c_str());
      else if (f.first == "grpc-message") message = PercentDecode(f.second);
      else if (f.first == ":status") http_status = f.second;
    }
    if (grpc_status >= 0) {
      call->status = grpc_status;
      call->status_message = message;
    }
    if (!http_status.empty() && http_status != "200" && grpc_status < 0) {
      call->status = kUnavailable;
      call->status_message = "HTTP status " + http_status;
    }
    if (end_stream) {
      if (call->status < 0) {
        call->status = kUnknown;
        call->status_message = "stream ended without a grpc-status";
      }
      if (!call->end_sent) {  // the server is done before we are: stop sending
        std::string code;
        h2::put_u32(&code, 0);
        out_ += h2::frame(h2::RST_STREAM, 0, call->stream_id, code);
      }
      Finish(call, call->status, call->status_message);
    }
  }

  void HandleData(uint32_t stream, const uint8_t* p, size_t n, size_t flow_bytes, bool end_stream) {
    conn_recv_consumed_ += static_cast<uint32_t>(flow_bytes);
    if (conn_recv_consumed_ >= kConnWindow / 2) {
      out_ += h2::window_update(0, conn_recv_consumed_);
      conn_recv_consumed_ = 0;
    }
    auto it = streams_.find(stream);
    if (it == streams_.end()) return;
    std::shared_ptr<GrpcCall> call = it->second;
    call->recv_consumed += static_cast<uint32_t>(flow_bytes);
    if (!end_stream && call->recv_consumed >= kStreamWindow / 2) {
      out_ += h2::window_update(stream, call->recv_consumed);
      call->recv_consumed = 0;
    }
    call->rx.append(reinterpret_cast<const char*>(p), n);
    size_t off = 0;
    while (call->rx.size() - off >= 5) {
      const uint8_t* h = reinterpret_cast<const uint8_t*>(call->rx.data()) + off;
      const uint32_t len = h2::get_u32(h + 1);
      if (call->rx.size() - off - 5 < len) break;
      std::string message = call->rx.substr(off + 5, len);
      if (h[0] != 0) {  // compressed with the response's grpc-encoding (deflate = zlib stream, or gzip)
        std::string plain;
        if (!detail::Inflate(message, &plain).IsOk()) {
          std::string code;
          h2::put_u32(&code, 8);
          out_ += h2::frame(h2::RST_STREAM, 0, stream, code);
          Finish(call, kInternal, "failed to decompress a response message");
          return;
        }
        message = std::move(plain);
      }
      off += 5 + static_cast<size_t>(len);
      if (call->on_message) call->on_message(std::move(message));
      else call->response = std::move(message);
    }
    if (off != 0) call->rx.erase(0, off);
    if (end_stream) Finish(call, call->status < 0 ? kUnknown : call->status, call->status < 0 ? "stream ended without trailers" : call->status_message);
  }

  bool HandleFrames() {
    size_t pos = 0;
    for (;;) {
      h2::FrameView f;
      const size_t n = h2::parse_frame(reinterpret_cast<const uint8_t*>(in_.data()) + pos, in_.size() - pos, &f);
      if (n == 0) break;
      pos += n;
      const uint8_t* p = f.payload;
      size_t len = f.length;
      switch (f.type) {
        case h2::SETTINGS:
          if (f.flags & h2::kAck) break;
          for (size_t i = 0; i + 6 <= len; i += 6) {
            const uint16_t id = static_cast<uint16_t>((p[i] << 8) | p[i + 1]);
            const uint32_t value = h2::get_u32(p + i + 2);
            if (id == h2::kSettingsInitialWindow) {
              const int64_t delta = static_cast<int64_t>(value) - peer_initial_window_;
              peer_initial_window_ = value;
              for (auto& kv : streams_) kv.second->send_window += delta;
            } else if (id == h2::kSettingsMaxFrame && value >= 16384) {
              peer_max_frame_ = value > (1u << 20) ? (1u << 20) : value;
            }
          }
          out_ += h2::frame(h2::SETTINGS, h2::kAck, 0, "");
          PumpAll();
          break;
        case h2::PING:
          if (!(f.flags & h2::kAck) && len == 8) out_ += h2::frame(h2::PING, h2::kAck, 0, std::string(reinterpret_cast<const char*>(p), 8));
          else if (f.flags & h2::kAck) ping_outstanding_ = false;
          break;
        case h2::WINDOW_UPDATE:
          if (len == 4) {
            const uint32_t inc = h2::get_u32(p) & 0x7FFFFFFFu;
            if (f.stream == 0) {
              conn_send_window_ += inc;
              PumpAll();
            } else {
              auto it = streams_.find(f.stream);
              if (it != streams_.end()) {
                it->second->send_window += inc;
                std::shared_ptr<GrpcCall> call = it->second;
                Pump(call);
              }
            }
          }
          break;
        case h2::HEADERS:
        case h2::CONTINUATION: {
          if (f.type == h2::HEADERS) {
            size_t pad = 0;
            if (f.flags & h2::kPadded) {
              if (len < 1) return false;
              pad = p[0];
              ++p;
              --len;
            }
            if (f.flags & h2::kPriority) {
              if (len < 5) return false;
              p += 5;
              len -= 5;
            }
            if (pad > len) return false;
            len -= pad;
            header_block_.assign(reinterpret_cast<const char*>(p), len);
            header_stream_ = f.stream;
            header_end_stream_ = (f.flags & h2::kEndStream) != 0;
          } else {
            header_block_.append(reinterpret_cast<const char*>(p), len);
          }
          if (f.flags & h2::kEndHeaders) {
            HandleHeaders(header_stream_, header_block_, header_end_stream_);
            header_block_.clear();
          }
          break;
        }
        case h2::DATA: {
          const size_t flow = len;
          if (f.flags & h2::kPadded) {
            if (len < 1 || p[0] > len - 1) return false;
            len -= 1 + p[0];
            ++p;
          }
          HandleData(f.stream, p, len, flow, (f.flags & h2::kEndStream) != 0);
          break;
        }
        case h2::RST_STREAM: {
          auto it = streams_.find(f.stream);
          if (it != streams_.end()) {
            std::shared_ptr<GrpcCall> call = it->second;
            const uint32_t code = len == 4 ? h2::get_u32(p) : 2;
            if (call->status >= 0 && code == 0) Finish(call, call->status, call->status_message);
            else Finish(call, code == 8 ? kCancelled : (code == 11 ? kResourceExhausted : kUnavailable), "stream reset by the server (HTTP/2 error " + std::to_string(code) + ")");
          }
          break;
        }
        case h2::GOAWAY:
          FailAll("the server closed the connection (GOAWAY)");
          break;
        default:
          break;
      }
    }
    if (pos) in_.erase(0, pos);
    return true;
  }

  void PumpAll() {
    std::vector<std::shared_ptr<GrpcCall>> calls;
    for (auto& kv : streams_) calls.push_back(kv.second);
    for (auto& c : calls) Pump(c);
  }

  void CheckDeadlines() {
    const Clock::time_point now = Clock::now();
    std::vector<std::shared_ptr<GrpcCall>> late;
    for (auto& kv : streams_) {
      if (kv.second->has_deadline && kv.second->deadline <= now) late.push_back(kv.second);
    }
    for (auto& call : late) {
      std::string code;
      h2::put_u32(&code, 8);
      out_ += h2::frame(h2::RST_STREAM, 0, call->stream_id, code);
      Finish(call, kDeadlineExceeded, "Deadline Exceeded");
    }
  }

  // GRPC keepalive (doc/keepalive.md): PING after keepalive_time of silence from the peer, the
  // connection is declared dead when the acknowledgement does not arrive within keepalive_timeout
  void KeepAliveTick() {
    if (keepalive_.time_ms == INT_MAX || broken_.load()) return;
    const Clock::time_point now = Clock::now();
    if (ping_outstanding_) {
      if (now - ping_sent_ > std::chrono::milliseconds(keepalive_.timeout_ms)) FailAll("keepalive watchdog timeout");
      return;
    }
    if ((streams_.empty() && !keepalive_.without_calls) || now - last_read_ < std::chrono::milliseconds(keepalive_.time_ms)) return;
    out_ += h2::frame(h2::PING, 0, 0, std::string("tb200png", 8));
    ping_outstanding_ = true;
    ping_sent_ = now;
  }

  int PollTimeoutMs() const {
    int ms = 200;
    if (keepalive_.time_ms != INT_MAX) ms = std::max(1, std::min(ms, std::min(keepalive_.time_ms, keepalive_.timeout_ms) / 2));
    const Clock::time_point now = Clock::now();
    for (const auto& kv : streams_) {
      if (!kv.second->has_deadline) continue;
      const auto left = std::chrono::duration_cast<std::chrono::microseconds>(kv.second->deadline - now).count();
      const int l = left <= 0 ? 0 : static_cast<int>((left + 999) / 1000);
      if (l < ms) ms = l;
    }
    return ms;
  }

  void IoMain() {
    char tmp[65536];
    for (;;) {
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (stop_) break;
      }
      HandleCommands();
      // flush
      while (!broken_.load() && out_off_ < out_.size()) {
        const ssize_t k = send(fd_, out_.data() + out_off_, out_.size() - out_off_, MSG_NOSIGNAL);
        if (k > 0) {
          out_off_ += static_cast<size_t>(k);
        } else if (k < 0 && errno == EINTR) {
          continue;
        } else if (k < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
          break;
        } else {
          FailAll("Socket closed");
          break;
        }
      }
      if (out_off_ == out_.size()) {
        out_.clear();
        out_off_ = 0;
      } else if (out_off_ > (1u << 20)) {
        out_.erase(0, out_off_);
        out_off_ = 0;
      }
      pollfd fds[2] = {{fd_, static_cast<short>(POLLIN | (out_off_ < out_.size() ? POLLOUT : 0)), 0}, {evfd_, POLLIN, 0}};
      if (broken_.load()) fds[0].events = 0;
      poll(fds, 2, PollTimeoutMs());
      if (fds[1].revents & POLLIN) {
        uint64_t count;
        if (read(evfd_, &count, sizeof(count)) < 0) count = 0;
      }
      if (!broken_.load() && (fds[0].revents & (POLLIN | POLLHUP | POLLERR))) {
        for (;;) {
          const ssize_t k = recv(fd_, tmp, sizeof(tmp), 0);
          if (k > 0) {
            in_.append(tmp, static_cast<size_t>(k));
            last_read_ = Clock::now();
            if (static_cast<size_t>(k) < sizeof(tmp)) break;
            continue;
          }
          if (k < 0 && errno == EINTR) continue;
          if (k < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) break;
          FailAll("Socket closed");
          break;
        }
        if (!broken_.load() && !HandleFrames()) FailAll("malformed HTTP/2 frame from the server");
      }
      CheckDeadlines();
      KeepAliveTick();
    }
    FailAll("the client was closed");
    // anything queued but never started
    std::deque<Command> cmds;
    {
      std::lock_guard<std::mutex> lk(mu_);
      cmds.swap(commands_);
    }
    for (Command& c : cmds) {
      if (c.kind == Command::START) Finish(c.call, kUnavailable, "the client was closed");
    }
  }

  int fd_, evfd_ = -1;
  std::string authority_;
  std::thread io_;
  std::mutex mu_;
  std::deque<Command> commands_;
  bool stop_ = false;
  std::atomic<bool> broken_{false};
  // I/O thread only
  std::string out_, in_;
  size_t out_off_ = 0;
  std::map<uint32_t, std::shared_ptr<GrpcCall>> streams_;
  uint32_t next_stream_id_ = 1;
  int64_t conn_send_window_ = h2::kDefaultWindow;
  int64_t peer_initial_window_ = h2::kDefaultWindow;
  uint32_t peer_max_frame_ = h2::kDefaultMaxFrame;
  uint32_t conn_recv_consumed_ = 0;
  KeepAlive keepalive_;
  bool ping_outstanding_ = false;
  Clock::time_point ping_sent_{}, last_read_ = Clock::now();
  h2::HpackDecoder hpack_;
  std::string header_block_;
  uint32_t header_stream_ = 0;
  bool header_end_stream_ = false;
};

}  // namespace detail

namespace {

// channels shared between clients of one URL (grpc_client.cc:60-150 keeps a similar map)
std::mutex g_cache_mu;
std::map<std::string, std::weak_ptr<detail::GrpcChannel>> g_cache;

constexpr char kService[] = "/inference.GRPCInferenceService/";

// ---- grpc_client.cc:178-446 -------------------------------------------------------------------
class InferResultGrpc : public InferResult {
 public:
  InferResultGrpc(std::shared_ptr<inference::ModelInferResponse> response, const Error& status)
      : response_(std::move(response)), request_status_(status) {
    Index();
  }
  explicit InferResultGrpc(std::shared_ptr<inference::ModelStreamInferResponse> stream_response)
      : stream_response_(std::move(stream_response)), request_status_(stream_response_->error_message()) {
    response_ = std::shared_ptr<inference::ModelInferResponse>(stream_response_, stream_response_->mutable_infer_response());
    Index();
  }
  Error ModelName(std::string* name) const override {
    *name = response_->model_name();
    return Error::Success;
  }
  Error ModelVersion(std::string* version) const override {
    *version = response_->model_version();
    return Error::Success;
  }
  Error Id(std::string* id) const override {
    *id = response_->id();
    return Error::Success;
  }
  Error Shape(const std::string& output_name, std::vector<int64_t>* shape) const override {
    shape->clear();
    const auto it = tensors_.find(output_name);
    if (it == tensors_.end()) return Error("The response does not contain shape for output name '" + output_name + "'");
    *shape = response_->outputs(it->second).shape();
    return Error::Success;
  }
  Error Datatype(const std::string& output_name, std::string* datatype) const override {
    const auto it = tensors_.find(output_name);
    if (it == tensors_.end()) return Error("The response does not contain datatype for output name '" + output_name + "'");
    *datatype = response_->outputs(it->second).datatype();
    return Error::Success;
  }
  Error RawData(const std::string& output_name, const uint8_t** buf, size_t* byte_size) const override {
    const auto it = tensors_.find(output_name);
    if (it == tensors_.end() || it->second >= response_->raw_output_contents_size()) {
      return Error("The response does not contain results for output name '" + output_name + "'");
    }
    const std::string& raw = response_->raw_output_contents(it->second);
    *buf = reinterpret_cast<const uint8_t*>(raw.data());
    *byte_size = raw.size();
    return Error::Success;
  }
  Error IsFinalResponse(bool* is_final_response) const override {
    if (is_final_response == nullptr) return Error("is_final_response cannot be nullptr");
    *is_final_response = is_final_response_;
    return Error::Success;
  }
  Error IsNullResponse(bool* is_null_response) const override {
    if (is_null_response == nullptr) return Error("is_null_response cannot be nullptr");
    *is_null_response = is_null_response_;
    return Error::Success;
  }
  Error StringData(const std::string& output_name, std::vector<std::string>* string_result) const override {
    std::string datatype;
    Error err = Datatype(output_name, &datatype);
    if (!err.IsOk()) return err;
    if (datatype != "BYTES") {
      return Error("This function supports tensors with datatype 'BYTES', requested output tensor '" + output_name +
                   "' with datatype '" + datatype + "'");
    }
    const uint8_t* buf;
    size_t byte_size;
    err = RawData(output_name, &buf, &byte_size);
    if (!err.IsOk()) return err;
    string_result->clear();
    size_t pos = 0;
    while (pos + 4 <= byte_size) {  // <u32 length><payload> per element
      uint32_t len;
      memcpy(&len, buf + pos, 4);
      pos += 4;
      if (len > byte_size - pos) break;
      string_result->emplace_back(reinterpret_cast<const char*>(buf + pos), len);
      pos += len;
    }
    return Error::Success;
  }
  std::string DebugString() const override { return response_->DebugString(); }
  Error RequestStatus() const override { return request_status_; }

 private:
  void Index() {
    for (int i = 0; i < response_->outputs_size(); ++i) tensors_[response_->outputs(i).name()] = i;
    const auto it = response_->parameters().find("triton_final_response");
    if (it != response_->parameters().end()) is_final_response_ = it->second.bool_param();
    is_null_response_ = response_->outputs_size() == 0 && is_final_response_;
  }
  std::shared_ptr<inference::ModelStreamInferResponse> stream_response_;
  std::shared_ptr<inference::ModelInferResponse> response_;
  Error request_status_;
  std::map<std::string, int> tensors_;
  bool is_final_response_ = true;
  bool is_null_response_ = false;
};

// request parameters / tensors (grpc_client.cc:1419-1580): everything except raw_input_contents
Error BuildInferHead(inference::ModelInferRequest* req, const InferOptions& options, const std::vector<InferInput*>& inputs,
                     const std::vector<const InferRequestedOutput*>& outputs) {
  req->set_model_name(options.model_name_);
  req->set_model_version(options.model_version_);
  req->set_id(options.request_id_);
  auto& params = *req->mutable_parameters();
  params["triton_enable_empty_final_response"].set_bool_param(options.triton_enable_empty_final_response_);
  if (options.sequence_id_ != 0 || !options.sequence_id_str_.empty()) {
    if (options.sequence_id_ != 0) params["sequence_id"].set_int64_param(static_cast<int64_t>(options.sequence_id_));
    else params["sequence_id"].set_string_param(options.sequence_id_str_);
    params["sequence_start"].set_bool_param(options.sequence_start_);
    params["sequence_end"].set_bool_param(options.sequence_end_);
  }
  if (options.priority_ != 0) params["priority"].set_uint64_param(options.priority_);
  if (options.server_timeout_ != 0) params["timeout"].set_int64_param(static_cast<int64_t>(options.server_timeout_));
  for (const auto& kv : options.request_parameters) {
    const RequestParameter& p = kv.second;
    if (p.type == "string") {
      params[kv.first].set_string_param(p.value);
    } else if (p.type == "int") {
      try {
        params[kv.first].set_int64_param(std::stoll(p.value));
      }
      catch (const std::exception&) {
        return Error("request parameter '" + kv.first + "' is not an integer: '" + p.value + "'");
      }
    } else if (p.type == "bool") {
      params[kv.first].set_bool_param(p.value == "true");
    }
  }
  for (InferInput* input : inputs) {
    auto* t = req->add_inputs();
    t->set_name(input->Name());
    t->set_datatype(input->Datatype());
    for (int64_t d : input->Shape()) t->add_shape(d);
    if (input->IsSharedMemory()) {
      std::string region;
      size_t byte_size = 0, offset = 0;
      input->SharedMemoryInfo(&region, &byte_size, &offset);
      auto& tp = *t->mutable_parameters();
      tp["shared_memory_region"].set_string_param(region);
      tp["shared_memory_byte_size"].set_int64_param(static_cast<int64_t>(byte_size));
      if (offset != 0) tp["shared_memory_offset"].set_int64_param(static_cast<int64_t>(offset));
    }
  }
  for (const InferRequestedOutput* output : outputs) {
    auto* t = req->add_outputs();
    t->set_name(output->Name());
    auto& tp = *t->mutable_parameters();
    if (output->ClassificationCount() != 0) tp["classification"].set_int64_param(static_cast<int64_t>(output->ClassificationCount()));
    if (output->IsSharedMemory()) {
      std::string region;
      size_t byte_size = 0, offset = 0;
      output->SharedMemoryInfo(&region, &byte_size, &offset);
      tp["shared_memory_region"].set_string_param(region);
      tp["shared_memory_byte_size"].set_int64_param(static_cast<int64_t>(byte_size));
      if (offset != 0) tp["shared_memory_offset"].set_int64_param(static_cast<int64_t>(offset));
    }
  }
  return Error::Success;
}

// head + raw_input_contents (field 7) straight from the scatter lists: one copy of the tensors
Error AppendInferMessage(std::string* message, const InferOptions& options, const std::vector<InferInput*>& inputs,
                         const std::vector<const InferRequestedOutput*>& outputs) {
  inference::ModelInferRequest head;
  Error err = BuildInferHead(&head, options, inputs, outputs);
  if (!err.IsOk()) return err;
  size_t total = 0;
  for (InferInput* input : inputs) {
    size_t n = 0;
    if (!input->IsSharedMemory()) input->ByteSize(&n);
    total += n + 11;
  }
  head.AppendTo(message);
  message->reserve(message->size() + total);
  for (InferInput* input : inputs) {
    if (input->IsSharedMemory()) continue;
    size_t n = 0;
    input->ByteSize(&n);
    input->PrepareForRequest();
    pb::put_tag(message, 7, pb::kBytes);
    pb::put_varint(message, n);
    bool end = false;
    size_t appended = 0;
    while (!end) {
      const uint8_t* buf = nullptr;
      size_t size = 0;
      input->GetNext(&buf, &size, &end);
      if (buf != nullptr && size != 0) {
        message->append(reinterpret_cast<const char*>(buf), size);
        appended += size;
      }
    }
    if (appended != n) return Error("input '" + input->Name() + "' changed size while the request was formed");
  }
  if (message->size() > static_cast<size_t>(INT_MAX)) {
    return Error("Request has byte size " + std::to_string(message->size()) + " which exceed gRPC's byte size limit " +
                 std::to_string(INT_MAX) + ".");
  }
  return Error::Success;
}

// message body -> zlib ("deflate") / gzip container, produced on the device (tb200_client.cc)
Error CompressOnDevice(std::string* body, grpc_compression_algorithm algorithm) {
  return detail::DeflateOnDevice(body, algorithm == GRPC_COMPRESS_GZIP);
}

}  // namespace

// =================================================================================================
InferenceServerGrpcClient::InferenceServerGrpcClient(const std::string& url, bool verbose, bool use_cached_channel,
                                                     const KeepAliveOptions& keepalive)
    : InferenceServerClient(verbose), keepalive_(keepalive), url_(url), use_cached_channel_(use_cached_channel) {
  worker_ = std::thread(&InferenceServerGrpcClient::CallbackWorker, this);
}

InferenceServerGrpcClient::~InferenceServerGrpcClient() {
  StopStream();
  {
    // nothing of this object may be touched by the I/O thread after this block
    std::shared_ptr<detail::GrpcChannel> channel;
    {
      std::lock_guard<std::mutex> lk(channel_mu_);
      channel = channel_;
    }
    std::unique_lock<std::mutex> lk(calls_mu_);
    if (channel) {
      for (auto& kv : active_calls_) channel->Cancel(kv.second);
    }
    calls_cv_.wait(lk, [this] { return active_calls_.empty(); });
  }
  {
    std::lock_guard<std::mutex> lk(worker_mu_);
    exiting_ = true;
  }
  worker_cv_.notify_all();
  if (worker_.joinable()) worker_.join();
  std::lock_guard<std::mutex> lk(channel_mu_);
  channel_.reset();
}

Error InferenceServerGrpcClient::Create(std::unique_ptr<InferenceServerGrpcClient>* client, const std::string& server_url,
                                        bool verbose, bool use_ssl, const SslOptions&, const KeepAliveOptions& keepalive_options,
                                        const bool use_cached_channel) {
  if (use_ssl) return Error("TLS is not built into this client: use_ssl must be false");
  client->reset(new InferenceServerGrpcClient(server_url, verbose, use_cached_channel, keepalive_options));
  return Error::Success;
}
Error InferenceServerGrpcClient::Create(std::unique_ptr<InferenceServerGrpcClient>* client, const std::string& server_url,
                                        const grpc::ChannelArguments& channel_args, bool verbose, bool use_ssl, const SslOptions&,
                                        const bool use_cached_channel) {
  if (use_ssl) return Error("TLS is not built into this client: use_ssl must be false");
  KeepAliveOptions keepalive;  // the keep-alive arguments among the generic ones
  const auto& ints = channel_args.ints();
  auto get = [&ints](const char* key, int* out) {
    const auto it = ints.find(key);
    if (it != ints.end()) *out = it->second;
  };
  int permit = keepalive.keepalive_permit_without_calls ? 1 : 0;
  get(GRPC_ARG_KEEPALIVE_TIME_MS, &keepalive.keepalive_time_ms);
  get(GRPC_ARG_KEEPALIVE_TIMEOUT_MS, &keepalive.keepalive_timeout_ms);
  get(GRPC_ARG_KEEPALIVE_PERMIT_WITHOUT_CALLS, &permit);
  get(GRPC_ARG_HTTP2_MAX_PINGS_WITHOUT_DATA, &keepalive.http2_max_pings_without_data);
  keepalive.keepalive_permit_without_calls = permit != 0;
  client->reset(new InferenceServerGrpcClient(server_url, verbose, use_cached_channel, keepalive));
  return Error::Success;
}

size_t InferenceServerGrpcClient::GetNumCachedChannels() const {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  size_t n = 0;
  for (const auto& kv : g_cache) {
    if (!kv.second.expired()) ++n;
  }
  return n;
}

Error InferenceServerGrpcClient::Channel(std::shared_ptr<detail::GrpcChannel>* channel) {
  std::lock_guard<std::mutex> lk(channel_mu_);
  if (channel_ && !channel_->Broken()) {
    *channel = channel_;
    return Error::Success;
  }
  channel_.reset();
  if (use_cached_channel_) {
    std::lock_guard<std::mutex> cache_lk(g_cache_mu);
    auto it = g_cache.find(url_);
    if (it != g_cache.end()) {
      std::shared_ptr<detail::GrpcChannel> cached = it->second.lock();
      if (cached && !cached->Broken()) channel_ = cached;
    }
  }
  if (!channel_) {
    detail::GrpcChannel::KeepAlive ka;
    ka.time_ms = keepalive_.keepalive_time_ms;
    ka.timeout_ms = keepalive_.keepalive_timeout_ms;
    ka.without_calls = keepalive_.keepalive_permit_without_calls;
    Error err = detail::GrpcChannel::Connect(url_, &channel_, ka);
    if (!err.IsOk()) return err;
    if (use_cached_channel_) {
      std::lock_guard<std::mutex> cache_lk(g_cache_mu);
      g_cache[url_] = channel_;
    }
  }
  *channel = channel_;
  return Error::Success;
}

Error InferenceServerGrpcClient::Unary(const char* method, const pb::Message& request, pb::Message* response,
                                       const Headers& headers, uint64_t timeout_us) {
  std::shared_ptr<detail::GrpcChannel> channel;
  Error err = Channel(&channel);
  if (!err.IsOk()) return err;
  auto call = std::make_shared<detail::GrpcCall>();
  call->path = std::string(kService) + method;
  call->metadata = headers;
  call->timeout_us = timeout_us;
  channel->Start(call, detail::GrpcChannel::Framed(request.SerializeAsString()), true);
  call->Wait();
  if (call->status != kOk) return Error(call->status_message.empty() ? "gRPC status " + std::to_string(call->status) : call->status_message);
  if (!response->ParseFromString(call->response)) return Error(std::string("malformed ") + method + " response");
  if (verbose_) std::cout << response->DebugString() << std::endl;
  return Error::Success;
}

Error InferenceServerGrpcClient::IsServerLive(bool* live, const Headers& headers, const uint64_t timeout_ms) {
  inference::ServerLiveRequest request;
  inference::ServerLiveResponse response;
  Error err = Unary("ServerLive", request, &response, headers, 1000 * timeout_ms);
  *live = err.IsOk() && response.live();
  if (verbose_ && err.IsOk()) std::cout << "Server Live : " << *live << std::endl;
  return err;
}
Error InferenceServerGrpcClient::IsServerReady(bool* ready, const Headers& headers, const uint64_t timeout_ms) {
  inference::ServerReadyRequest request;
  inference::ServerReadyResponse response;
  Error err = Unary("ServerReady", request, &response, headers, 1000 * timeout_ms);
  *ready = err.IsOk() && response.ready();
  if (verbose_ && err.IsOk()) std::cout << "Server Ready : " << *ready << std::endl;
  return err;
}
Error InferenceServerGrpcClient::IsModelReady(bool* ready, const std::string& model_name, const std::string& model_version,
                                              const Headers& headers, const uint64_t timeout_ms) {
  inference::ModelReadyRequest request;
  request.set_name(model_name);
  request.set_version(model_version);
  inference::ModelReadyResponse response;
  Error err = Unary("ModelReady", request, &response, headers, 1000 * timeout_ms);
  *ready = err.IsOk() && response.ready();
  if (verbose_ && err.IsOk()) {
    std::cout << "Model Ready : name: " << model_name;
    if (!model_version.empty()) std::cout << "(version: " << model_version << ") ";
    std::cout << ": " << *ready << std::endl;
  }
  return err;
}
Error InferenceServerGrpcClient::ServerMetadata(inference::ServerMetadataResponse* server_metadata, const Headers& headers, const uint64_t timeout_ms) {
  inference::ServerMetadataRequest request;
  return Unary("ServerMetadata", request, server_metadata, headers, 1000 * timeout_ms);
}
Error InferenceServerGrpcClient::ModelMetadata(inference::ModelMetadataResponse* model_metadata, const std::string& model_name,
                                               const std::string& model_version, const Headers& headers, const uint64_t timeout_ms) {
  inference::ModelMetadataRequest request;
  request.set_name(model_name);
  request.set_version(model_version);
  return Unary("ModelMetadata", request, model_metadata, headers, 1000 * timeout_ms);
}
Error InferenceServerGrpcClient::ModelConfig(inference::ModelConfigResponse* model_config, const std::string& model_name,
                                             const std::string& model_version, const Headers& headers, const uint64_t timeout_ms) {
  inference::ModelConfigRequest request;
  request.set_name(model_name);
  request.set_version(model_version);
  return Unary("ModelConfig", request, model_config, headers, 1000 * timeout_ms);
}
Error InferenceServerGrpcClient::ModelRepositoryIndex(inference::RepositoryIndexResponse* repository_index, const Headers& headers, const uint64_t timeout_ms) {
  inference::RepositoryIndexRequest request;
  return Unary("RepositoryIndex", request, repository_index, headers, 1000 * timeout_ms);
}
Error InferenceServerGrpcClient::LoadModel(const std::string& model_name, const Headers& headers, const std::string& config,
                                           const std::map<std::string, std::vector<char>>& files, const uint64_t timeout_ms) {
  inference::RepositoryModelLoadRequest request;
  request.set_model_name(model_name);
  if (!config.empty()) (*request.mutable_parameters())["config"].set_string_param(config);
  for (const auto& kv : files) (*request.mutable_parameters())[kv.first].set_bytes_param(kv.second.data(), kv.second.size());
  inference::RepositoryModelLoadResponse response;
  Error err = Unary("RepositoryModelLoad", request, &response, headers, 1000 * timeout_ms);
  if (verbose_ && err.IsOk()) std::cout << "Loaded model '" << model_name << "'" << std::endl;
  return err;
}
Error InferenceServerGrpcClient::UnloadModel(const std::string& model_name, const Headers& headers, const uint64_t timeout_ms) {
  inference::RepositoryModelUnloadRequest request;
  request.set_model_name(model_name);
  inference::RepositoryModelUnloadResponse response;
  Error err = Unary("RepositoryModelUnload", request, &response, headers, 1000 * timeout_ms);
  if (verbose_ && err.IsOk()) std::cout << "Unloaded model '" << model_name << "'" << std::endl;
  return err;
}
Error InferenceServerGrpcClient::ModelInferenceStatistics(inference::ModelStatisticsResponse* infer_stat, const std::string& model_name,
                                                          const std::string& model_version, const Headers& headers, const uint64_t timeout_ms) {
  inference::ModelStatisticsRequest request;
  request.set_name(model_name);
  request.set_version(model_version);
  return Unary("ModelStatistics", request, infer_stat, headers, 1000 * timeout_ms);
}
Error InferenceServerGrpcClient::UpdateTraceSettings(inference::TraceSettingResponse* response, const std::string& model_name,
                                                     const std::map<std::string, std::vector<std::string>>& settings,
                                                     const Headers& headers, const uint64_t timeout_ms) {
  inference::TraceSettingRequest request;
  if (!model_name.empty()) request.set_model_name(model_name);
  for (const auto& kv : settings) {
    auto& value = (*request.mutable_settings())[kv.first];  // an empty list clears the setting
    for (const std::string& v : kv.second) value.add_value(v);
  }
  return Unary("TraceSetting", request, response, headers, 1000 * timeout_ms);
}
Error InferenceServerGrpcClient::GetTraceSettings(inference::TraceSettingResponse* settings, const std::string& model_name,
                                                  const Headers& headers, const uint64_t timeout_ms) {
  inference::TraceSettingRequest request;
  if (!model_name.empty()) request.set_model_name(model_name);
  return Unary("TraceSetting", request, settings, headers, 1000 * timeout_ms);
}
Error InferenceServerGrpcClient::SystemSharedMemoryStatus(inference::SystemSharedMemoryStatusResponse* status,
                                                          const std::string& region_name, const Headers& headers, const uint64_t timeout_ms) {
  inference::SystemSharedMemoryStatusRequest request;
  request.set_name(region_name);
  return Unary("SystemSharedMemoryStatus", request, status, headers, 1000 * timeout_ms);
}
Error InferenceServerGrpcClient::RegisterSystemSharedMemory(const std::string& name, const std::string& key, const size_t byte_size,
                                                            const size_t offset, const Headers& headers, const uint64_t timeout_ms) {
  inference::SystemSharedMemoryRegisterRequest request;
  request.set_name(name);
  request.set_key(key);
  request.set_offset(offset);
  request.set_byte_size(byte_size);
  inference::SystemSharedMemoryRegisterResponse response;
  Error err = Unary("SystemSharedMemoryRegister", request, &response, headers, 1000 * timeout_ms);
  if (verbose_ && err.IsOk()) std::cout << "Registered system shared memory with name '" << name << "'" << std::endl;
  return err;
}
Error InferenceServerGrpcClient::UnregisterSystemSharedMemory(const std::string& name, const Headers& headers, const uint64_t timeout_ms) {
  inference::SystemSharedMemoryUnregisterRequest request;
  request.set_name(name);
  inference::SystemSharedMemoryUnregisterResponse response;
  Error err = Unary("SystemSharedMemoryUnregister", request, &response, headers, 1000 * timeout_ms);
  if (verbose_ && err.IsOk()) {
    if (!name.empty()) std::cout << "Unregistered system shared memory with name '" << name << "'" << std::endl;
    else std::cout << "Unregistered all system shared memory regions" << std::endl;
  }
  return err;
}
Error InferenceServerGrpcClient::CudaSharedMemoryStatus(inference::CudaSharedMemoryStatusResponse* status,
                                                        const std::string& region_name, const Headers& headers, const uint64_t timeout_ms) {
  inference::CudaSharedMemoryStatusRequest request;
  request.set_name(region_name);
  return Unary("CudaSharedMemoryStatus", request, status, headers, 1000 * timeout_ms);
}
Error InferenceServerGrpcClient::RegisterCudaSharedMemoryRaw(const std::string& name, const uint8_t* handle64, const size_t device_id,
                                                             const size_t byte_size, const Headers& headers, const uint64_t timeout_ms) {
  inference::CudaSharedMemoryRegisterRequest request;
  request.set_name(name);
  request.set_raw_handle(handle64, 64);
  request.set_device_id(static_cast<int64_t>(device_id));
  request.set_byte_size(byte_size);
  inference::CudaSharedMemoryRegisterResponse response;
  Error err = Unary("CudaSharedMemoryRegister", request, &response, headers, 1000 * timeout_ms);
  if (verbose_ && err.IsOk()) std::cout << "Registered cuda shared memory with name '" << name << "'" << std::endl;
  return err;
}
Error InferenceServerGrpcClient::UnregisterCudaSharedMemory(const std::string& name, const Headers& headers, const uint64_t timeout_ms) {
  inference::CudaSharedMemoryUnregisterRequest request;
  request.set_name(name);
  inference::CudaSharedMemoryUnregisterResponse response;
  Error err = Unary("CudaSharedMemoryUnregister", request, &response, headers, 1000 * timeout_ms);
  if (verbose_ && err.IsOk()) {
    if (!name.empty()) std::cout << "Unregistered cuda shared memory with name '" << name << "'" << std::endl;
    else std::cout << "Unregistered all cuda shared memory regions" << std::endl;
  }
  return err;
}

// ---- inference ----------------------------------------------------------------------------------
Error InferenceServerGrpcClient::SerializeInferRequest(std::string* message, const InferOptions& options,
                                                       const std::vector<InferInput*>& inputs,
                                                       const std::vector<const InferRequestedOutput*>& outputs) {
  message->clear();
  return AppendInferMessage(message, options, inputs, outputs);
}

Error InferenceServerGrpcClient::StartInfer(std::shared_ptr<detail::GrpcCall>* out, const InferOptions& options,
                                            const std::vector<InferInput*>& inputs,
                                            const std::vector<const InferRequestedOutput*>& outputs, const Headers& headers,
                                            grpc_compression_algorithm compression_algorithm,
                                            std::function<void(detail::GrpcCall*)> on_done, RequestTimers* timer) {
  std::shared_ptr<detail::GrpcChannel> channel;
  Error err = Channel(&channel);
  if (!err.IsOk()) return err;
  std::string framed(5, '\0');
  err = AppendInferMessage(&framed, options, inputs, outputs);
  if (!err.IsOk()) return err;
  auto call = std::make_shared<detail::GrpcCall>();
  if (compression_algorithm != GRPC_COMPRESS_NONE) {
    std::string body = framed.substr(5);
    err = CompressOnDevice(&body, compression_algorithm);
    if (!err.IsOk()) return err;
    framed = detail::GrpcChannel::Framed(body, true);
    call->encoding = compression_algorithm == GRPC_COMPRESS_GZIP ? "gzip" : "deflate";
  } else {
    h2::put_grpc_prefix(reinterpret_cast<uint8_t*>(&framed[0]), static_cast<uint32_t>(framed.size() - 5));
  }
  call->path = std::string(kService) + "ModelInfer";
  call->metadata = headers;
  call->timeout_us = options.client_timeout_;
  call->on_done = std::move(on_done);
  // the send timer covers the marshalling (as in the reference); it must be complete before the
  // call can finish on the I/O thread
  if (timer != nullptr) timer->CaptureTimestamp(RequestTimers::Kind::SEND_END);
  if (call->on_done) {  // asynchronous: the destructor cancels and waits for it
    std::lock_guard<std::mutex> lk(calls_mu_);
    active_calls_[call.get()] = call;
  }
  channel->Start(call, std::move(framed), true);
  *out = std::move(call);
  return Error::Success;
}

namespace {
InferResult* MakeResult(detail::GrpcCall* call, bool verbose) {
  auto response = std::make_shared<inference::ModelInferResponse>();
  Error status;
  if (call->status != kOk) {
    status = Error(call->status_message.empty() ? "gRPC status " + std::to_string(call->status) : call->status_message);
  } else if (!response->ParseFromString(call->response)) {
    status = Error("malformed ModelInfer response");
  } else if (verbose) {
    std::cout << response->DebugString() << std::endl;
  }
  return new InferResultGrpc(std::move(response), status);
}
}  // namespace

Error InferenceServerGrpcClient::Infer(InferResult** result, const InferOptions& options, const std::vector<InferInput*>& inputs,
                                       const std::vector<const InferRequestedOutput*>& outputs, const Headers& headers,
                                       grpc_compression_algorithm compression_algorithm) {
  RequestTimers timer;
  timer.CaptureTimestamp(RequestTimers::Kind::REQUEST_START);
  timer.CaptureTimestamp(RequestTimers::Kind::SEND_START);  // marshalling, as in the reference
  std::shared_ptr<detail::GrpcCall> call;
  Error err = StartInfer(&call, options, inputs, outputs, headers, compression_algorithm, nullptr, &timer);
  if (!err.IsOk()) return err;
  call->Wait();
  timer.CaptureTimestamp(RequestTimers::Kind::RECV_START);
  *result = MakeResult(call.get(), verbose_);
  timer.CaptureTimestamp(RequestTimers::Kind::RECV_END);
  timer.CaptureTimestamp(RequestTimers::Kind::REQUEST_END);
  err = UpdateInferStat(timer);
  if (!err.IsOk()) std::cerr << "Failed to update context stat: " << err << std::endl;
  return (*result)->RequestStatus();
}

void InferenceServerGrpcClient::Dispatch(std::function<void()> fn) {
  // notified under the lock: the destructor joins the worker, which cannot run (and finish) the
  // job before this thread is done with the members
  std::lock_guard<std::mutex> lk(worker_mu_);
  worker_jobs_.push_back(std::move(fn));
  worker_cv_.notify_one();
}

void InferenceServerGrpcClient::CallbackWorker() {
  for (;;) {
    std::function<void()> job;
    {
      std::unique_lock<std::mutex> lk(worker_mu_);
      worker_cv_.wait(lk, [this] { return exiting_ || !worker_jobs_.empty(); });
      if (worker_jobs_.empty()) return;  // exiting and drained
      job = std::move(worker_jobs_.front());
      worker_jobs_.pop_front();
    }
    job();
  }
}

Error InferenceServerGrpcClient::AsyncInfer(OnCompleteFn callback, const InferOptions& options, const std::vector<InferInput*>& inputs,
                                            const std::vector<const InferRequestedOutput*>& outputs, const Headers& headers,
                                            grpc_compression_algorithm compression_algorithm) {
  if (callback == nullptr) return Error("Callback function must be provided along with AsyncInfer() call.");
  auto timer = std::make_shared<RequestTimers>();
  timer->CaptureTimestamp(RequestTimers::Kind::REQUEST_START);
  timer->CaptureTimestamp(RequestTimers::Kind::SEND_START);
  std::shared_ptr<detail::GrpcCall> call;
  Error err = StartInfer(&call, options, inputs, outputs, headers, compression_algorithm,
                         [this, callback, timer](detail::GrpcCall* done) {
                           // I/O thread: hand the finished call to the callback thread, then
                           // drop it from the calls the destructor waits for (last access to `this`)
                           struct Deregister {
                             InferenceServerGrpcClient* self;
                             detail::GrpcCall* call;
                             ~Deregister() {
                               std::lock_guard<std::mutex> lk(self->calls_mu_);
                               self->active_calls_.erase(call);
                               self->calls_cv_.notify_all();
                             }
                           } deregister{this, done};
                           auto status = done->status;
                           auto message = std::make_shared<std::string>(std::move(done->response));
                           auto status_message = done->status_message;
                           Dispatch([this, callback, timer, status, message, status_message] {
                             detail::GrpcCall finished;
                             finished.status = status;
                             finished.status_message = status_message;
                             finished.response = std::move(*message);
                             timer->CaptureTimestamp(RequestTimers::Kind::RECV_START);
                             InferResult* result = MakeResult(&finished, verbose_);
                             timer->CaptureTimestamp(RequestTimers::Kind::RECV_END);
                             timer->CaptureTimestamp(RequestTimers::Kind::REQUEST_END);
                             Error stat = UpdateInferStat(*timer);
                             if (!stat.IsOk()) std::cerr << "Failed to update context stat: " << stat << std::endl;
                             callback(result);
                           });
                         },
                         timer.get());
  return err;
}

Error InferenceServerGrpcClient::InferMulti(std::vector<InferResult*>* results, const std::vector<InferOptions>& options,
                                            const std::vector<std::vector<InferInput*>>& inputs,
                                            const std::vector<std::vector<const InferRequestedOutput*>>& outputs,
                                            const Headers& headers, grpc_compression_algorithm compression_algorithm) {
  if (options.size() != 1 && options.size() != inputs.size()) {
    return Error("'options' must either contain 1 element or match size of 'inputs'");
  }
  if (outputs.size() > 1 && outputs.size() != inputs.size()) {
    return Error("'outputs' must either contain 0/1 element or match size of 'inputs'");
  }
  static const std::vector<const InferRequestedOutput*> kNoOutputs;
  for (size_t i = 0; i < inputs.size(); ++i) {
    const InferOptions& opt = options.size() == 1 ? options[0] : options[i];
    const auto& outs = outputs.empty() ? kNoOutputs : (outputs.size() == 1 ? outputs[0] : outputs[i]);
    InferResult* result = nullptr;
    Error err = Infer(&result, opt, inputs[i], outs, headers, compression_algorithm);
    if (result != nullptr) results->push_back(result);
    if (!err.IsOk()) return err;
  }
  return Error::Success;
}

Error InferenceServerGrpcClient::AsyncInferMulti(OnMultiCompleteFn callback, const std::vector<InferOptions>& options,
                                                 const std::vector<std::vector<InferInput*>>& inputs,
                                                 const std::vector<std::vector<const InferRequestedOutput*>>& outputs,
                                                 const Headers& headers, grpc_compression_algorithm compression_algorithm) {
  if (options.size() != 1 && options.size() != inputs.size()) {
    return Error("'options' must either contain 1 element or match size of 'inputs'");
  }
  if (outputs.size() > 1 && outputs.size() != inputs.size()) {
    return Error("'outputs' must either contain 0/1 element or match size of 'inputs'");
  }
  if (callback == nullptr) return Error("Callback function must be provided along with AsyncInferMulti() call.");
  struct Gather {
    std::mutex mu;
    std::vector<InferResult*> results;
    size_t left;
  };
  auto gather = std::make_shared<Gather>();
  gather->results.resize(inputs.size(), nullptr);
  gather->left = inputs.size();
  static const std::vector<const InferRequestedOutput*> kNoOutputs;
  for (size_t i = 0; i < inputs.size(); ++i) {
    const InferOptions& opt = options.size() == 1 ? options[0] : options[i];
    const auto& outs = outputs.empty() ? kNoOutputs : (outputs.size() == 1 ? outputs[0] : outputs[i]);
    Error err = AsyncInfer(
        [gather, callback, i](InferResult* result) {
          bool last;
          {
            std::lock_guard<std::mutex> lk(gather->mu);
            gather->results[i] = result;
            last = --gather->left == 0;
          }
          if (last) callback(gather->results);
        },
        opt, inputs[i], outs, headers, compression_algorithm);
    if (!err.IsOk()) return err;
  }
  return Error::Success;
}

// ---- streaming ------------------------------------------------------------------------------------
Error InferenceServerGrpcClient::StartStream(OnCompleteFn callback, bool enable_stats, uint32_t stream_timeout, const Headers& headers,
                                             grpc_compression_algorithm compression_algorithm) {
  if (stream_call_) {
    return Error(
        "cannot start another stream with one already running. 'InferenceServerClient' supports only a single active stream at a "
        "given time.");
  }
  if (callback == nullptr) return Error("Callback function must be provided along with StartStream() call.");
  if (compression_algorithm != GRPC_COMPRESS_NONE) return Error("stream compression is not supported by this client");
  std::shared_ptr<detail::GrpcChannel> channel;
  Error err = Channel(&channel);
  if (!err.IsOk()) return err;
  stream_callback_ = callback;
  enable_stream_stats_ = enable_stats;
  auto call = std::make_shared<detail::GrpcCall>();
  call->path = std::string(kService) + "ModelStreamInfer";
  call->metadata = headers;
  call->timeout_us = stream_timeout;
  call->on_message = [this](std::string&& bytes) {
    auto raw = std::make_shared<std::string>(std::move(bytes));
    Dispatch([this, raw] {
      std::unique_ptr<RequestTimers> timer;
      if (enable_stream_stats_) {
        std::lock_guard<std::mutex> lk(stream_mu_);
        if (!ongoing_stream_request_timers_.empty()) {
          timer = std::move(ongoing_stream_request_timers_.front());
          ongoing_stream_request_timers_.pop();
        }
      }
      if (timer) timer->CaptureTimestamp(RequestTimers::Kind::RECV_START);
      auto response = std::make_shared<inference::ModelStreamInferResponse>();
      if (!response->ParseFromString(*raw)) response->set_error_message("malformed ModelStreamInfer response");
      if (verbose_) std::cout << response->DebugString() << std::endl;
      InferResult* result = new InferResultGrpc(std::move(response));
      if (timer) {
        timer->CaptureTimestamp(RequestTimers::Kind::RECV_END);
        timer->CaptureTimestamp(RequestTimers::Kind::REQUEST_END);
        Error stat = UpdateInferStat(*timer);
        if (!stat.IsOk()) std::cerr << "Failed to update context stat: " << stat << std::endl;
      }
      stream_callback_(result);
    });
  };
  call->on_done = [this](detail::GrpcCall* done) {
    const int status = done->status;
    const std::string message = done->status_message;
    Dispatch([this, status, message] {
      if (status != kOk && status != kCancelled) {
        // the call ended abnormally: the user hears about it through the callback, once
        auto response = std::make_shared<inference::ModelStreamInferResponse>();
        response->set_error_message(message.empty() ? "gRPC status " + std::to_string(status) : message);
        stream_callback_(new InferResultGrpc(std::move(response)));
      }
      {
        std::lock_guard<std::mutex> lk(stream_mu_);
        stream_done_ = true;
      }
      stream_cv_.notify_all();
    });
  };
  {
    std::lock_guard<std::mutex> lk(stream_mu_);
    stream_done_ = false;
  }
  stream_call_ = call;
  channel->Start(call, std::string(), false);
  if (verbose_) std::cout << "Started stream..." << std::endl;
  return Error::Success;
}

Error InferenceServerGrpcClient::StopStream() {
  if (!stream_call_) return Error::Success;
  std::shared_ptr<detail::GrpcChannel> channel;
  {
    std::lock_guard<std::mutex> lk(channel_mu_);
    channel = channel_;
  }
  if (channel) channel->WritesDone(stream_call_);
  {
    // the responses still in flight are delivered before this returns
    std::unique_lock<std::mutex> lk(stream_mu_);
    stream_cv_.wait(lk, [this] { return stream_done_; });
    while (!ongoing_stream_request_timers_.empty()) ongoing_stream_request_timers_.pop();
  }
  stream_call_.reset();
  if (verbose_) std::cout << "Stopped stream..." << std::endl;
  return Error::Success;
}

Error InferenceServerGrpcClient::AsyncStreamInfer(const InferOptions& options, const std::vector<InferInput*>& inputs,
                                                  const std::vector<const InferRequestedOutput*>& outputs) {
  if (!stream_call_) return Error("Stream has been closed.");
  std::unique_ptr<RequestTimers> timer;
  if (enable_stream_stats_) {
    timer.reset(new RequestTimers());
    timer->CaptureTimestamp(RequestTimers::Kind::REQUEST_START);
    timer->CaptureTimestamp(RequestTimers::Kind::SEND_START);
  }
  std::string framed(5, '\0');
  Error err = AppendInferMessage(&framed, options, inputs, outputs);
  if (!err.IsOk()) return err;
  h2::put_grpc_prefix(reinterpret_cast<uint8_t*>(&framed[0]), static_cast<uint32_t>(framed.size() - 5));
  {
    std::lock_guard<std::mutex> lk(stream_mu_);
    if (stream_done_) return Error("Stream has been closed.");
    if (timer) {
      timer->CaptureTimestamp(RequestTimers::Kind::SEND_END);
      ongoing_stream_request_timers_.push(std::move(timer));
    }
  }
  std::shared_ptr<detail::GrpcChannel> channel;
  {
    std::lock_guard<std::mutex> lk(channel_mu_);
    channel = channel_;
  }
  if (!channel || channel->Broken()) return Error("Stream has been closed.");
  channel->Write(stream_call_, std::move(framed));
  if (verbose_) {
    std::cout << "Sent request";
    if (!options.request_id_.empty()) std::cout << " '" << options.request_id_ << "'";
    std::cout << " to the stream" << std::endl;
  }
  return Error::Success;
}

}}  // namespace tb200::client

// =================================================================================================
// C entry points over the channel for other front ends (the Python drop-in's opt-in native gRPC
// transport, client_b200/grpc/_native_channel.py): one blocking unary call, bytes in, bytes out.
// Declared in tb200_grpc_channel.h.
#include "tb200_grpc_channel.h"

struct tb200c_grpc_channel {
  std::string url;
  std::mutex mu;
  std::shared_ptr<tb200::client::detail::GrpcChannel> channel;
};

extern "C" {

int tb200c_grpc_channel_open(const char* url, tb200c_grpc_channel** out) {
  if (url == nullptr || out == nullptr) return -1;
  tb200c_grpc_channel* c = new tb200c_grpc_channel();
  c->url = url;
  *out = c;
  return 0;  // the connection is made by the first call (and re-made after a failure)
}

void tb200c_grpc_channel_close(tb200c_grpc_channel* c) { delete c; }

int tb200c_grpc_unary(tb200c_grpc_channel* c, const char* path, const uint8_t* request, uint64_t request_bytes, const char* const* metadata,
                      int metadata_pairs, uint64_t timeout_us, uint8_t** response, uint64_t* response_bytes, char* message, uint64_t message_cap) {
  using namespace tb200::client;
  auto fail = [&](int status, const std::string& text) {
    if (message != nullptr && message_cap != 0) snprintf(message, message_cap, "%s", text.c_str());
    if (response != nullptr) *response = nullptr;
    if (response_bytes != nullptr) *response_bytes = 0;
    return status;
  };
  if (c == nullptr || path == nullptr || response == nullptr || response_bytes == nullptr || (request == nullptr && request_bytes != 0)) {
    return fail(3, "NULL argument");
  }
  std::shared_ptr<detail::GrpcChannel> channel;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->channel || c->channel->Broken()) {
      c->channel.reset();
      Error err = detail::GrpcChannel::Connect(c->url, &c->channel);
      if (!err.IsOk()) return fail(kUnavailable, err.Message());
    }
    channel = c->channel;
  }
  auto call = std::make_shared<detail::GrpcCall>();
  call->path = path;
  call->timeout_us = timeout_us;
  for (int i = 0; i < metadata_pairs; ++i) call->metadata[metadata[2 * i]] = metadata[2 * i + 1];
  std::string framed(5, '\0');
  framed.append(reinterpret_cast<const char*>(request), request_bytes);
  tb200::h2::put_grpc_prefix(reinterpret_cast<uint8_t*>(&framed[0]), static_cast<uint32_t>(request_bytes));
  channel->Start(call, std::move(framed), true);
  call->Wait();
  if (call->status != kOk) return fail(call->status, call->status_message);
  *response_bytes = call->response.size();
  *response = static_cast<uint8_t*>(malloc(call->response.size() == 0 ? 1 : call->response.size()));
  if (*response == nullptr) return fail(kResourceExhausted, "out of memory");
  memcpy(*response, call->response.data(), call->response.size());
  if (message != nullptr && message_cap != 0) message[0] = '\0';
  return 0;
}

void tb200c_free(void* p) { free(p); }

}  // extern "C"
