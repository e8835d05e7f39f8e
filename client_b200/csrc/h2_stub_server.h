// h2_stub_server.h -- canned-response gRPC server (cleartext HTTP/2, prior knowledge) for
// measuring the native load generator's gRPC transport.  TOOLING (the reference has no
// server).  Every unary call gets: HEADERS(:status 200, content-type
// application/grpc), DATA(5-byte prefix + the canned message), HEADERS(grpc-status 0,
// END_STREAM).  Request header blocks are not decoded; request DATA is only counted for flow
// control.  A few epoll threads own the connections.
#ifndef TB200_CSRC_H2_STUB_SERVER_H_
#define TB200_CSRC_H2_STUB_SERVER_H_

#include <arpa/inet.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "h2.h"

namespace tb200 {

class H2StubServer {
 public:
  ~H2StubServer() { Stop(); }

  // Stream mode (ModelStreamInfer): every request MESSAGE on a stream is answered with
  // `responses_per_request - 1` copies of `message` and one `final_message`; the response
  // HEADERS go out once per stream, the trailers when the client half-closes.
  bool StartStreaming(const char* host, int* port, int nthreads, const std::string& message, const std::string& final_message,
                      int responses_per_request) {
    streaming_ = true;
    auto framed = [](const std::string& m) {
      std::string d(5, '\0');
      h2::put_grpc_prefix(reinterpret_cast<uint8_t*>(&d[0]), static_cast<uint32_t>(m.size()));
      return h2::frame(h2::DATA, 0, 0, d + m);
    };
    stream_reply_.clear();
    for (int i = 0; i + 1 < responses_per_request; ++i) {
      reply_id_off_.push_back(stream_reply_.size() + 5);
      stream_reply_ += framed(message);
    }
    reply_id_off_.push_back(stream_reply_.size() + 5);
    stream_reply_ += framed(final_message);
    return Start(host, port, nthreads, message);
  }

  bool Start(const char* host, int* port, int nthreads, const std::string& message) {
    // response image of one call; the three stream ids are patched per call
    const std::string hdr = h2::grpc_response_headers(), trl = h2::grpc_trailers_ok();
    std::string data(5, '\0');
    h2::put_grpc_prefix(reinterpret_cast<uint8_t*>(&data[0]), static_cast<uint32_t>(message.size()));
    data += message;
    response_ = h2::frame(h2::HEADERS, h2::kEndHeaders, 0, hdr);
    id_off_[0] = 5;
    id_off_[1] = response_.size() + 5;
    response_ += h2::frame(h2::DATA, 0, 0, data);
    id_off_[2] = response_.size() + 5;
    response_ += h2::frame(h2::HEADERS, h2::kEndHeaders | h2::kEndStream, 0, trl);
    hello_ = h2::frame(h2::SETTINGS, 0, 0, h2::setting(h2::kSettingsInitialWindow, 1u << 20)) +
             h2::window_update(0, (1u << 30) - h2::kDefaultWindow);

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
    for (int i = 0; i < (nthreads < 1 ? 1 : nthreads); ++i) {
      std::unique_ptr<Loop> l(new Loop());
      l->epfd = epoll_create1(0);
      l->evfd = eventfd(0, EFD_NONBLOCK);
      epoll_event ev{};
      ev.events = EPOLLIN;
      ev.data.u32 = kEvTag;
      epoll_ctl(l->epfd, EPOLL_CTL_ADD, l->evfd, &ev);
      loops_.push_back(std::move(l));
    }
    for (size_t i = 0; i < loops_.size(); ++i) threads_.emplace_back(&H2StubServer::LoopMain, this, loops_[i].get());
    acceptor_ = std::thread(&H2StubServer::AcceptMain, this);
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

  uint64_t calls() const { return calls_.load(); }

 private:
  static constexpr uint32_t kEvTag = 0xFFFFFFFFu;
  struct StreamState {  // stream mode: where we are inside the client's message sequence
    bool headers_sent = false;
    uint8_t prefix[5];
    uint32_t prefix_have = 0;
    uint64_t body_left = 0;
    uint64_t consumed = 0;
  };
  struct Conn {
    int fd = -1;
    bool preface = false;
    uint64_t consumed = 0;
    std::string buf;
    std::map<uint32_t, StreamState> streams;
  };
  struct Loop {
    int epfd = -1, evfd = -1;
    std::mutex mu;
    std::vector<int> new_fds;
    std::vector<Conn> conns;
  };

  static void Kick(Loop* l) {
    const uint64_t one = 1;
    if (write(l->evfd, &one, sizeof(one)) < 0) return;
  }
  static bool SendAll(int fd, const std::string& data) {
    size_t off = 0;
    while (off < data.size()) {
      const ssize_t k = send(fd, data.data() + off, data.size() - off, MSG_NOSIGNAL | MSG_DONTWAIT);
      if (k > 0) {
        off += static_cast<size_t>(k);
      } else if (k < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
        pollfd p{fd, POLLOUT, 0};
        if (poll(&p, 1, 1000) <= 0) return false;
      } else if (k < 0 && errno == EINTR) {
        continue;
      } else {
        return false;
      }
    }
    return true;
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
    c.fd = -1;
    c.preface = false;
    c.consumed = 0;
    c.buf.clear();
    c.streams.clear();
  }

  static void PatchIds(std::string* out, size_t base, const std::vector<size_t>& offsets, uint32_t stream) {
    for (size_t off : offsets) {
      (*out)[base + off + 0] = static_cast<char>((stream >> 24) & 0x7F);
      (*out)[base + off + 1] = static_cast<char>(stream >> 16);
      (*out)[base + off + 2] = static_cast<char>(stream >> 8);
      (*out)[base + off + 3] = static_cast<char>(stream);
    }
  }

  // stream mode: DATA payload of one stream; a reply per completed request message
  void ServeStreamData(Conn& c, const h2::FrameView& f, std::string* out) {
    StreamState& st = c.streams[f.stream];
    if (!st.headers_sent) {
      *out += h2::frame(h2::HEADERS, h2::kEndHeaders, f.stream, h2::grpc_response_headers());
      st.headers_sent = true;
    }
    const uint8_t* p = f.payload;
    size_t left = f.length;
    while (left > 0) {
      if (st.body_left == 0 && st.prefix_have < 5) {
        const size_t n = std::min<size_t>(5 - st.prefix_have, left);
        memcpy(st.prefix + st.prefix_have, p, n);
        st.prefix_have += static_cast<uint32_t>(n);
        p += n;
        left -= n;
        if (st.prefix_have < 5) break;
        st.body_left = h2::get_u32(st.prefix + 1);
      }
      const size_t n = static_cast<size_t>(std::min<uint64_t>(st.body_left, left));
      st.body_left -= n;
      p += n;
      left -= n;
      if (st.body_left == 0 && st.prefix_have == 5) {  // one request message complete
        st.prefix_have = 0;
        const size_t base = out->size();
        *out += stream_reply_;
        PatchIds(out, base, reply_id_off_, f.stream);
        calls_.fetch_add(1, std::memory_order_relaxed);
      }
    }
    st.consumed += f.length;
    if (st.consumed >= (1u << 18) && !(f.flags & h2::kEndStream)) {
      *out += h2::window_update(f.stream, static_cast<uint32_t>(st.consumed));
      st.consumed = 0;
    }
    if (f.flags & h2::kEndStream) {
      *out += h2::frame(h2::HEADERS, h2::kEndHeaders | h2::kEndStream, f.stream, h2::grpc_trailers_ok());
      c.streams.erase(f.stream);
    }
  }

  // everything complete in c.buf; false when the connection should be dropped
  bool Serve(Conn& c) {
    static const char kPreface[] = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n";
    std::string out;
    size_t pos = 0;
    if (!c.preface) {
      if (c.buf.size() < 24) return true;
      if (memcmp(c.buf.data(), kPreface, 24) != 0) return false;
      c.preface = true;
      pos = 24;
      out += hello_;
    }
    for (;;) {
      h2::FrameView f;
      const size_t n = h2::parse_frame(reinterpret_cast<const uint8_t*>(c.buf.data()) + pos, c.buf.size() - pos, &f);
      if (n == 0) break;
      pos += n;
      bool answer = false;
      switch (f.type) {
        case h2::SETTINGS:
          if (!(f.flags & h2::kAck)) out += h2::frame(h2::SETTINGS, h2::kAck, 0, "");
          break;
        case h2::PING:
          if (!(f.flags & h2::kAck) && f.length == 8) out += h2::frame(h2::PING, h2::kAck, 0, std::string(reinterpret_cast<const char*>(f.payload), 8));
          break;
        case h2::DATA:
          c.consumed += f.length;
          if (streaming_) ServeStreamData(c, f, &out);
          else answer = (f.flags & h2::kEndStream) != 0;
          break;
        case h2::HEADERS:
          answer = !streaming_ && (f.flags & h2::kEndStream) != 0;  // a call without a message
          break;
        case h2::RST_STREAM:
          c.streams.erase(f.stream);
          break;
        default:
          break;
      }
      if (answer && f.stream != 0) {
        const size_t base = out.size();
        out += response_;
        PatchIds(&out, base, id_off_, f.stream);
        calls_.fetch_add(1, std::memory_order_relaxed);
      }
    }
    if (pos) c.buf.erase(0, pos);
    if (c.consumed >= (1u << 28)) {
      out += h2::window_update(0, static_cast<uint32_t>(c.consumed));
      c.consumed = 0;
    }
    return out.empty() || SendAll(c.fd, out);
  }

  void LoopMain(Loop* l) {
    epoll_event events[128];
    std::vector<int> fds;
    char tmp[65536];
    while (!stop_.load(std::memory_order_relaxed)) {
      const int n = epoll_wait(l->epfd, events, 128, 100);
      for (int e = 0; e < n; ++e) {
        const uint32_t tag = events[e].data.u32;
        if (tag == kEvTag) {
          uint64_t count;
          if (read(l->evfd, &count, sizeof(count)) < 0) continue;
          {
            std::lock_guard<std::mutex> lk(l->mu);
            fds.swap(l->new_fds);
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
          continue;
        }
        if (tag >= l->conns.size() || l->conns[tag].fd < 0) continue;
        Conn& c = l->conns[tag];
        bool closed = false;
        for (;;) {
          const ssize_t k = recv(c.fd, tmp, sizeof(tmp), MSG_DONTWAIT);
          if (k > 0) {
            c.buf.append(tmp, static_cast<size_t>(k));
            if (static_cast<size_t>(k) < sizeof(tmp)) break;
            continue;
          }
          if (k < 0 && errno == EINTR) continue;
          if (k < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) break;
          closed = true;
          break;
        }
        if (closed || !Serve(c)) CloseConn(l, c);
      }
    }
  }

  std::string response_, hello_;
  std::vector<size_t> id_off_{0, 0, 0};
  bool streaming_ = false;
  std::string stream_reply_;          // DATA frames answering one request message (stream mode)
  std::vector<size_t> reply_id_off_;  // where their stream ids are
  int listen_fd_ = -1;
  std::atomic<bool> stop_{false};
  std::atomic<uint64_t> calls_{0};
  std::thread acceptor_;
  std::vector<std::unique_ptr<Loop>> loops_;
  std::vector<std::thread> threads_;
};

}  // namespace tb200

#endif  // TB200_CSRC_H2_STUB_SERVER_H_
