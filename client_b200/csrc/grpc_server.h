// grpc_server.h -- a small epoll-driven gRPC server core (cleartext HTTP/2, prior knowledge) for
// the native stand-in model server (mock_server.cu) and the echo server the CPU tests use.
// TOOLING: the reference has no server.
//
// Same shape as http_server.h: a few event-loop threads own many connections each; a handler
// answers a message at once or defers, the answer then comes from any thread through
// CompleteLater() + Flush() (one eventfd write per event-loop thread and batch).  Both unary calls
// and bidirectional streams are "a sequence of request messages on an HTTP/2 stream": the handler
// sees every message with the stream's :path and says with each reply whether the call is over.
// HPACK: request header blocks are decoded (h2.h, dynamic table included) for :path only;
// response blocks are literal-only.  Flow control is honoured in both directions.
#ifndef TB200_CSRC_GRPC_SERVER_H_
#define TB200_CSRC_GRPC_SERVER_H_

#include <arpa/inet.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "h2.h"

namespace tb200 {

struct GrpcReply {
  std::vector<std::string> messages;  // response messages (unframed protobuf bytes), sent in order
  bool finish = true;                 // send the trailers after them (unary: always; stream: at half-close)
  int status = 0;                     // grpc-status of the trailers
  std::string status_message;         // grpc-message
  // optional: fills in the fields above on the event-loop thread that owns the connection, right
  // before the reply is queued -- response serialisation then runs on all loop threads instead of
  // the one thread that called CompleteLater()
  std::function<void(GrpcReply*)> build;
};

class EpollGrpcServer {
 public:
  // One request message of the call `call_id` on `path`.  `half_close`: the client has ended its
  // side (true for every unary call; on a stream it comes as a last call with an empty message and
  // `is_message` false).  Return true with *reply filled to answer now, false to defer: then
  // CompleteLater(call_id, ...) must follow.
  using Handler = std::function<bool(uint64_t call_id, const std::string& path, std::string&& message, bool is_message,
                                     bool half_close, GrpcReply* reply)>;

  ~EpollGrpcServer() { Stop(); }

  bool Start(const char* host, int* port, int nthreads, Handler handler) {
    handler_ = std::move(handler);
    listen_fd_ = socket(AF_INET, SOCK_STREAM, 0);
    int one = 1;
    setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    addr.sin_port = htons(static_cast<uint16_t>(*port));
    if (listen_fd_ < 0 || inet_pton(AF_INET, host, &addr.sin_addr) != 1 ||
        bind(listen_fd_, reinterpret_cast<sockaddr*>(&addr), sizeof(addr)) != 0 || listen(listen_fd_, 1024) != 0) {
      if (listen_fd_ >= 0) close(listen_fd_);
      listen_fd_ = -1;
      return false;
    }
    socklen_t len = sizeof(addr);
    getsockname(listen_fd_, reinterpret_cast<sockaddr*>(&addr), &len);
    *port = ntohs(addr.sin_port);
    for (int i = 0; i < std::max(1, std::min(nthreads, 60)); ++i) {
      std::unique_ptr<Loop> l(new Loop());
      l->epfd = epoll_create1(0);
      l->evfd = eventfd(0, EFD_NONBLOCK);
      epoll_event ev{};
      ev.events = EPOLLIN;
      ev.data.u32 = kEvTag;
      epoll_ctl(l->epfd, EPOLL_CTL_ADD, l->evfd, &ev);
      loops_.push_back(std::move(l));
    }
    for (size_t i = 0; i < loops_.size(); ++i) threads_.emplace_back(&EpollGrpcServer::LoopMain, this, static_cast<uint32_t>(i));
    acceptor_ = std::thread(&EpollGrpcServer::AcceptMain, this);
    return true;
  }

  void Stop() {
    if (listen_fd_ < 0) return;
    stop_.store(true);
    shutdown(listen_fd_, SHUT_RDWR);
    close(listen_fd_);
    listen_fd_ = -1;
    if (acceptor_.joinable()) acceptor_.join();
    for (auto& l : loops_) Kick(l.get());
    for (std::thread& t : threads_) {
      if (t.joinable()) t.join();
    }
    threads_.clear();
    for (auto& l : loops_) {
      for (Conn& c : l->conns) {
        if (c.fd >= 0) close(c.fd);
      }
      close(l->evfd);
      close(l->epfd);
    }
    loops_.clear();
  }

  // queue the deferred answer of `call_id`; nothing is sent before Flush()
  void CompleteLater(uint64_t call_id, GrpcReply&& reply) {
    Loop* l = loops_[static_cast<size_t>(call_id >> 58)].get();
    std::lock_guard<std::mutex> lk(l->mu);
    l->done.push_back(Done{call_id, std::move(reply)});
    l->touched = true;
  }
  void Flush() {
    for (auto& l : loops_) {
      bool kick;
      {
        std::lock_guard<std::mutex> lk(l->mu);
        kick = l->touched;
        l->touched = false;
      }
      if (kick) Kick(l.get());
    }
  }

 private:
  static constexpr uint32_t kEvTag = 0xFFFFFFFFu;
  static constexpr uint32_t kRecvStreamWindow = 4u << 20;
  static constexpr uint32_t kRecvConnWindow = 1u << 30;
  // call id: loop (6 bits) | connection index (14) | connection generation (12) | stream id (31)
  static uint64_t CallId(uint32_t loop, uint32_t index, uint32_t gen, uint32_t stream) {
    return (static_cast<uint64_t>(loop) << 58) | (static_cast<uint64_t>(index & 0x3FFF) << 44) |
           (static_cast<uint64_t>(gen & 0xFFF) << 32) | (stream & 0x7FFFFFFFu);
  }
  struct Stream {
    std::string path;
    std::string rx;                    // request bytes not yet cut into messages
    std::deque<std::string> pending;   // DATA payload bytes (framed messages) waiting for window
    size_t pending_off = 0;
    int64_t send_window = 0;
    uint32_t recv_consumed = 0;
    bool headers_sent = false, half_closed = false, want_trailers = false, trailers_sent = false;
    int status = 0;
    std::string status_message;
  };
  struct Conn {
    int fd = -1;
    uint32_t gen = 0;
    bool preface = false, want_out = false;
    std::string in, out;
    size_t out_off = 0;
    h2::HpackDecoder hpack;
    std::map<uint32_t, Stream> streams;
    int64_t conn_send_window = h2::kDefaultWindow;
    int64_t peer_initial_window = h2::kDefaultWindow;
    uint32_t peer_max_frame = h2::kDefaultMaxFrame;
    uint32_t conn_recv_consumed = 0;
    std::string header_block;
    uint32_t header_stream = 0;
    bool header_end_stream = false;
  };
  struct Done {
    uint64_t call_id;
    GrpcReply reply;
  };
  struct Loop {
    int epfd = -1, evfd = -1;
    std::mutex mu;
    std::vector<int> new_fds;
    std::vector<Done> done;
    bool touched = false;
    std::vector<Conn> conns;
  };

  static void Kick(Loop* l) {
    const uint64_t one = 1;
    if (write(l->evfd, &one, sizeof(one)) < 0) return;
  }

  void AcceptMain() {
    size_t next = 0;
    while (!stop_.load()) {
      const int fd = accept(listen_fd_, nullptr, nullptr);
      if (fd < 0) break;
      int one = 1;
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
      fcntl(fd, F_SETFL, fcntl(fd, F_GETFL, 0) | O_NONBLOCK);
      Loop* l = loops_[next++ % loops_.size()].get();
      {
        std::lock_guard<std::mutex> lk(l->mu);
        l->new_fds.push_back(fd);
      }
      Kick(l);
    }
  }

  void CloseConn(Loop* l, Conn& c) {
    epoll_ctl(l->epfd, EPOLL_CTL_DEL, c.fd, nullptr);
    close(c.fd);
    const uint32_t gen = c.gen + 1;
    c = Conn();
    c.gen = gen;
  }

  // ---- sending ------------------------------------------------------------------------------
  void QueueReply(Conn& c, uint32_t stream_id, GrpcReply&& reply) {
    auto it = c.streams.find(stream_id);
    if (it == c.streams.end()) return;  // reset by the client meanwhile
    Stream& st = it->second;
    if (!st.headers_sent && !(reply.messages.empty() && reply.finish)) {
      c.out += h2::frame(h2::HEADERS, h2::kEndHeaders, stream_id, h2::grpc_response_headers());
      st.headers_sent = true;
    }
    for (std::string& m : reply.messages) {
      std::string framed(5, '\0');
      h2::put_grpc_prefix(reinterpret_cast<uint8_t*>(&framed[0]), static_cast<uint32_t>(m.size()));
      framed += m;
      st.pending.push_back(std::move(framed));
    }
    if (reply.finish) {
      st.want_trailers = true;
      st.status = reply.status;
      st.status_message = std::move(reply.status_message);
    }
    Pump(c, stream_id);
  }

  static std::string PercentEncode(const std::string& s) {
    static const char* hex = "0123456789ABCDEF";
    std::string out;
    for (unsigned char ch : s) {
      if (ch >= 0x20 && ch < 0x7F && ch != '%') {
        out.push_back(static_cast<char>(ch));
      } else {
        out.push_back('%');
        out.push_back(hex[ch >> 4]);
        out.push_back(hex[ch & 15]);
      }
    }
    return out;
  }

  void Pump(Conn& c, uint32_t stream_id) {
    auto it = c.streams.find(stream_id);
    if (it == c.streams.end()) return;
    Stream& st = it->second;
    while (!st.pending.empty()) {
      const std::string& m = st.pending.front();
      const size_t left = m.size() - st.pending_off;
      const int64_t room = std::min<int64_t>(st.send_window, c.conn_send_window);
      if (room <= 0) return;
      const size_t n = std::min<size_t>(std::min<size_t>(left, static_cast<size_t>(room)), c.peer_max_frame);
      uint8_t hdr[9];
      h2::put_frame_header(hdr, static_cast<uint32_t>(n), h2::DATA, 0, stream_id);
      c.out.append(reinterpret_cast<const char*>(hdr), 9);
      c.out.append(m, st.pending_off, n);
      st.send_window -= static_cast<int64_t>(n);
      c.conn_send_window -= static_cast<int64_t>(n);
      if (n == left) {
        st.pending.pop_front();
        st.pending_off = 0;
      } else {
        st.pending_off += n;
      }
    }
    if (st.want_trailers && !st.trailers_sent) {
      std::string block;
      if (!st.headers_sent) {  // trailers-only response: :status and content-type come with them
        block = h2::grpc_response_headers();
        st.headers_sent = true;
      }
      h2::hpack_literal(&block, "grpc-status", std::to_string(st.status));
      if (!st.status_message.empty()) h2::hpack_literal(&block, "grpc-message", PercentEncode(st.status_message));
      c.out += h2::frame(h2::HEADERS, h2::kEndHeaders | h2::kEndStream, stream_id, block);
      st.trailers_sent = true;
      if (!st.half_closed) {  // we are done before the client is: tell it to stop (RFC 9113 8.1)
        std::string code;
        h2::put_u32(&code, 0);
        c.out += h2::frame(h2::RST_STREAM, 0, stream_id, code);
      }
      c.streams.erase(it);
    }
  }

  bool FlushOut(Loop* l, Conn& c, uint32_t index) {
    while (c.out_off < c.out.size()) {
      const ssize_t k = send(c.fd, c.out.data() + c.out_off, c.out.size() - c.out_off, MSG_NOSIGNAL | MSG_DONTWAIT);
      if (k > 0) {
        c.out_off += static_cast<size_t>(k);
      } else if (k < 0 && errno == EINTR) {
        continue;
      } else if (k < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
        break;
      } else {
        return false;
      }
    }
    const bool want = c.out_off < c.out.size();
    if (!want) {
      c.out.clear();
      c.out_off = 0;
    } else if (c.out_off > (1u << 20)) {
      c.out.erase(0, c.out_off);
      c.out_off = 0;
    }
    if (want != c.want_out) {
      epoll_event ev{};
      ev.events = EPOLLIN | (want ? EPOLLOUT : 0);
      ev.data.u32 = index;
      epoll_ctl(l->epfd, EPOLL_CTL_MOD, c.fd, &ev);
      c.want_out = want;
    }
    return true;
  }

  // ---- receiving ----------------------------------------------------------------------------
  void Dispatch(uint32_t loop, Conn& c, uint32_t index, uint32_t stream_id, std::string&& message, bool is_message, bool half_close) {
    auto it = c.streams.find(stream_id);
    if (it == c.streams.end()) return;
    GrpcReply reply;
    const uint64_t id = CallId(loop, index, c.gen, stream_id);
    if (handler_(id, it->second.path, std::move(message), is_message, half_close, &reply)) QueueReply(c, stream_id, std::move(reply));
  }

  void OnData(uint32_t loop, Conn& c, uint32_t index, const h2::FrameView& f, const uint8_t* p, size_t n) {
    c.conn_recv_consumed += f.length;
    if (c.conn_recv_consumed >= kRecvConnWindow / 2) {
      c.out += h2::window_update(0, c.conn_recv_consumed);
      c.conn_recv_consumed = 0;
    }
    auto it = c.streams.find(f.stream);
    if (it == c.streams.end()) return;
    const bool end = (f.flags & h2::kEndStream) != 0;
    {
      Stream& st = it->second;
      st.recv_consumed += f.length;
      if (!end && st.recv_consumed >= kRecvStreamWindow / 2) {
        c.out += h2::window_update(f.stream, st.recv_consumed);
        st.recv_consumed = 0;
      }
      st.rx.append(reinterpret_cast<const char*>(p), n);
      if (end) st.half_closed = true;
    }
    // cut complete messages; the handler may erase the stream (QueueReply with finish)
    bool delivered_close = false;
    for (;;) {
      it = c.streams.find(f.stream);
      if (it == c.streams.end()) return;
      Stream& st = it->second;
      if (st.rx.size() < 5) break;
      const uint8_t* h = reinterpret_cast<const uint8_t*>(st.rx.data());
      const size_t len = h2::get_u32(h + 1);
      if (st.rx.size() - 5 < len) break;
      if (h[0] != 0) {  // compressed request messages: we advertise identity only
        GrpcReply r;
        r.status = 12;
        r.status_message = "compressed request messages are not supported by this server";
        QueueReply(c, f.stream, std::move(r));
        return;
      }
      std::string message = st.rx.substr(5, len);
      st.rx.erase(0, 5 + len);
      const bool last = end && st.rx.empty();
      delivered_close = delivered_close || last;
      Dispatch(loop, c, index, f.stream, std::move(message), true, last);
    }
    if (end && !delivered_close) Dispatch(loop, c, index, f.stream, std::string(), false, true);
  }

  void OnHeaders(uint32_t loop, Conn& c, uint32_t index) {
    std::vector<h2::HpackDecoder::Field> fields;
    if (!c.hpack.Decode(reinterpret_cast<const uint8_t*>(c.header_block.data()), c.header_block.size(), &fields)) {
      c.out += h2::frame(h2::GOAWAY, 0, 0, std::string("\0\0\0\0\0\0\0\x09", 8));  // COMPRESSION_ERROR
      return;
    }
    if (c.streams.count(c.header_stream)) return;  // trailers from a client: nothing to do
    Stream st;
    for (const auto& f : fields) {
      if (f.first == ":path") st.path = f.second;
    }
    st.send_window = c.peer_initial_window;
    st.half_closed = c.header_end_stream;
    c.streams.emplace(c.header_stream, std::move(st));
    if (c.header_end_stream) Dispatch(loop, c, index, c.header_stream, std::string(), false, true);
  }

  bool OnReadable(uint32_t loop, Conn& c, uint32_t index) {
    static const char kPreface[] = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n";
    size_t pos = 0;
    if (!c.preface) {
      if (c.in.size() < 24) return true;
      if (memcmp(c.in.data(), kPreface, 24) != 0) return false;
      c.preface = true;
      pos = 24;
      c.out += h2::frame(h2::SETTINGS, 0, 0, h2::setting(h2::kSettingsInitialWindow, kRecvStreamWindow));
      c.out += h2::window_update(0, kRecvConnWindow - h2::kDefaultWindow);
    }
    for (;;) {
      h2::FrameView f;
      const size_t n = h2::parse_frame(reinterpret_cast<const uint8_t*>(c.in.data()) + pos, c.in.size() - pos, &f);
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
              const int64_t delta = static_cast<int64_t>(value) - c.peer_initial_window;
              c.peer_initial_window = value;
              for (auto& kv : c.streams) kv.second.send_window += delta;
            } else if (id == h2::kSettingsMaxFrame && value >= 16384) {
              c.peer_max_frame = std::min<uint32_t>(value, 1u << 20);
            }
          }
          c.out += h2::frame(h2::SETTINGS, h2::kAck, 0, "");
          PumpAll(c);
          break;
        case h2::PING:
          if (!(f.flags & h2::kAck) && len == 8) c.out += h2::frame(h2::PING, h2::kAck, 0, std::string(reinterpret_cast<const char*>(p), 8));
          break;
        case h2::WINDOW_UPDATE:
          if (len == 4) {
            const uint32_t inc = h2::get_u32(p) & 0x7FFFFFFFu;
            if (f.stream == 0) {
              c.conn_send_window += inc;
              PumpAll(c);
            } else {
              auto it = c.streams.find(f.stream);
              if (it != c.streams.end()) {
                it->second.send_window += inc;
                Pump(c, f.stream);
              }
            }
          }
          break;
        case h2::HEADERS:
        case h2::CONTINUATION:
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
            c.header_block.assign(reinterpret_cast<const char*>(p), len);
            c.header_stream = f.stream;
            c.header_end_stream = (f.flags & h2::kEndStream) != 0;
          } else {
            c.header_block.append(reinterpret_cast<const char*>(p), len);
          }
          if (f.flags & h2::kEndHeaders) {
            OnHeaders(loop, c, index);
            c.header_block.clear();
          }
          break;
        case h2::DATA: {
          if (f.flags & h2::kPadded) {
            if (len < 1 || p[0] > len - 1) return false;
            len -= 1 + p[0];
            ++p;
          }
          OnData(loop, c, index, f, p, len);
          break;
        }
        case h2::RST_STREAM:
          c.streams.erase(f.stream);
          break;
        case h2::GOAWAY:
          break;
        default:
          break;
      }
    }
    if (pos) c.in.erase(0, pos);
    return true;
  }

  void PumpAll(Conn& c) {
    std::vector<uint32_t> ids;
    for (auto& kv : c.streams) ids.push_back(kv.first);
    for (uint32_t id : ids) Pump(c, id);
  }

  void LoopMain(uint32_t loop) {
    Loop* l = loops_[loop].get();
    epoll_event events[128];
    std::vector<int> fds;
    std::vector<Done> done;
    std::vector<uint32_t> dirty;
    char tmp[65536];
    while (!stop_.load(std::memory_order_relaxed)) {
      const int n = epoll_wait(l->epfd, events, 128, 100);
      dirty.clear();
      for (int e = 0; e < n; ++e) {
        const uint32_t tag = events[e].data.u32;
        if (tag == kEvTag) {
          uint64_t count;
          if (read(l->evfd, &count, sizeof(count)) < 0) continue;
          {
            std::lock_guard<std::mutex> lk(l->mu);
            fds.swap(l->new_fds);
            done.swap(l->done);
          }
          for (int fd : fds) {
            uint32_t index = 0;
            while (index < l->conns.size() && l->conns[index].fd >= 0) ++index;
            if (index == l->conns.size()) l->conns.emplace_back();
            l->conns[index].fd = fd;
            epoll_event ev{};
            ev.events = EPOLLIN;
            ev.data.u32 = index;
            epoll_ctl(l->epfd, EPOLL_CTL_ADD, fd, &ev);
          }
          fds.clear();
          for (Done& d : done) {
            if (d.reply.build) {
              std::function<void(GrpcReply*)> build = std::move(d.reply.build);
              d.reply.build = nullptr;
              build(&d.reply);  // runs even when the connection is gone: it may own resources to release
            }
            const uint32_t index = static_cast<uint32_t>((d.call_id >> 44) & 0x3FFF);
            if (index >= l->conns.size()) continue;
            Conn& c = l->conns[index];
            if (c.fd < 0 || (c.gen & 0xFFF) != ((d.call_id >> 32) & 0xFFF)) continue;  // that connection is gone
            QueueReply(c, static_cast<uint32_t>(d.call_id & 0x7FFFFFFFu), std::move(d.reply));
            dirty.push_back(index);
          }
          done.clear();
          continue;
        }
        if (tag >= l->conns.size() || l->conns[tag].fd < 0) continue;
        Conn& c = l->conns[tag];
        bool closed = false;
        if (events[e].events & (EPOLLIN | EPOLLHUP | EPOLLERR)) {
          for (;;) {
            const ssize_t k = recv(c.fd, tmp, sizeof(tmp), MSG_DONTWAIT);
            if (k > 0) {
              c.in.append(tmp, static_cast<size_t>(k));
              if (static_cast<size_t>(k) < sizeof(tmp)) break;
              continue;
            }
            if (k < 0 && errno == EINTR) continue;
            if (k < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) break;
            closed = true;
            break;
          }
          if (!closed && !OnReadable(loop, c, tag)) closed = true;
        }
        if (closed) CloseConn(l, c);
        else dirty.push_back(tag);
      }
      for (uint32_t index : dirty) {
        Conn& c = l->conns[index];
        if (c.fd >= 0 && !FlushOut(l, c, index)) CloseConn(l, c);
      }
    }
  }

  Handler handler_;
  int listen_fd_ = -1;
  std::atomic<bool> stop_{false};
  std::thread acceptor_;
  std::vector<std::unique_ptr<Loop>> loops_;
  std::vector<std::thread> threads_;
};

}  // namespace tb200

#endif  // TB200_CSRC_GRPC_SERVER_H_
