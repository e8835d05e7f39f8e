// http_server.h -- a small epoll-driven HTTP/1.1 server core for the two stand-in servers
// (canned-response stub in loadgen.cc, CUDA-shared-memory model server in mock_server.cu).
// TOOLING: the reference has no server.
//
// A few event-loop threads own many keep-alive connections each.  A handler either answers
// at once or defers: the answer then comes later from any thread through CompleteLater() +
// Flush(), which costs one eventfd write per event-loop thread and batch -- so a model pass
// that served 200 requests wakes 16 threads, not 200.
#ifndef TB200_CSRC_HTTP_SERVER_H_
#define TB200_CSRC_HTTP_SERVER_H_

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
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace tb200 {

struct HttpRequest {
  std::string method;
  std::string path;  // without the query string
  std::string body;
};

class EpollHttpServer {
 public:
  // Return true with *response filled (complete HTTP response bytes) to answer now; return
  // false to defer -- then CompleteLater(conn_id, ...) must follow eventually.
  using Handler = std::function<bool(uint64_t conn_id, const HttpRequest& req, std::string* response)>;

  ~EpollHttpServer() { Stop(); }

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
    for (int i = 0; i < std::max(1, nthreads); ++i) {
      std::unique_ptr<Loop> l(new Loop());
      l->epfd = epoll_create1(0);
      l->evfd = eventfd(0, EFD_NONBLOCK);
      epoll_event ev{};
      ev.events = EPOLLIN;
      ev.data.u32 = kEvTag;
      epoll_ctl(l->epfd, EPOLL_CTL_ADD, l->evfd, &ev);
      loops_.push_back(std::move(l));
    }
    for (size_t i = 0; i < loops_.size(); ++i) threads_.emplace_back(&EpollHttpServer::LoopMain, this, static_cast<uint32_t>(i));
    acceptor_ = std::thread(&EpollHttpServer::AcceptMain, this);
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

  // queue the deferred answer of `conn_id`; nothing is sent before Flush()
  void CompleteLater(uint64_t conn_id, const std::string& response) {
    Loop* l = loops_[static_cast<size_t>(conn_id >> 48)].get();
    std::lock_guard<std::mutex> lk(l->mu);
    l->done.push_back(Done{conn_id, response});
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

  static std::string Response(int status, const std::string& body) {
    return "HTTP/1.1 " + std::to_string(status) + (status == 200 ? " OK" : " Bad Request") +
           "\r\nContent-Type: application/json\r\nContent-Length: " + std::to_string(body.size()) + "\r\n\r\n" + body;
  }

 private:
  static constexpr uint32_t kEvTag = 0xFFFFFFFFu;
  struct Conn {
    int fd = -1;
    uint32_t gen = 0;
    bool busy = false;  // a deferred request is outstanding
    std::string buf;
  };
  struct Done {
    uint64_t conn_id;
    std::string response;
  };
  struct Loop {
    int epfd = -1, evfd = -1;
    std::mutex mu;
    std::vector<int> new_fds;
    std::vector<Done> done;
    bool touched = false;
    std::vector<Conn> conns;  // touched by the loop thread only
  };

  static void Kick(Loop* l) {
    const uint64_t one = 1;
    if (write(l->evfd, &one, sizeof(one)) < 0) return;
  }
  // conn id: loop (16 bits) | generation (16 bits) | index (32 bits)
  static uint64_t MakeId(uint32_t loop, uint32_t gen, uint32_t index) {
    return (static_cast<uint64_t>(loop) << 48) | (static_cast<uint64_t>(gen & 0xFFFFu) << 32) | index;
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

  static void SendAll(int fd, const std::string& data) {
    size_t off = 0;
    while (off < data.size()) {
      const ssize_t k = send(fd, data.data() + off, data.size() - off, MSG_NOSIGNAL | MSG_DONTWAIT);
      if (k > 0) {
        off += static_cast<size_t>(k);
      } else if (k < 0 && (errno == EAGAIN || errno == EWOULDBLOCK)) {
        pollfd p{fd, POLLOUT, 0};
        if (poll(&p, 1, 1000) <= 0) return;
      } else if (k < 0 && errno == EINTR) {
        continue;
      } else {
        return;
      }
    }
  }

  void CloseConn(Loop* l, Conn& c) {
    epoll_ctl(l->epfd, EPOLL_CTL_DEL, c.fd, nullptr);
    close(c.fd);
    c.fd = -1;
    c.gen += 1;
    c.busy = false;
    c.buf.clear();
  }

  // every complete request sitting in c.buf (one at a time: the next is held while a
  // deferred answer is outstanding)
  void Serve(uint32_t li, Loop* l, uint32_t index) {
    Conn& c = l->conns[index];
    while (c.fd >= 0 && !c.busy) {
      const size_t he = c.buf.find("\r\n\r\n");
      if (he == std::string::npos) return;
      const size_t header_end = he + 4;
      size_t clen = 0;
      for (size_t i = 0; i + 15 <= header_end; ++i) {
        if ((i == 0 || c.buf[i - 1] == '\n') && strncasecmp(c.buf.c_str() + i, "content-length:", 15) == 0) {
          clen = strtoull(c.buf.c_str() + i + 15, nullptr, 10);
          break;
        }
      }
      if (c.buf.size() < header_end + clen) return;
      HttpRequest req;
      const size_t sp1 = c.buf.find(' ');
      const size_t sp2 = c.buf.find(' ', sp1 + 1);
      if (sp1 == std::string::npos || sp2 == std::string::npos || sp2 > header_end) {
        CloseConn(l, c);
        return;
      }
      req.method = c.buf.substr(0, sp1);
      req.path = c.buf.substr(sp1 + 1, sp2 - sp1 - 1);
      const size_t q = req.path.find('?');
      if (q != std::string::npos) req.path.resize(q);
      req.body = c.buf.substr(header_end, clen);
      c.buf.erase(0, header_end + clen);
      std::string response;
      const uint64_t id = MakeId(li, c.gen, index);
      if (handler_(id, req, &response)) {
        SendAll(c.fd, response);
      } else {
        c.busy = true;
      }
    }
  }

  void LoopMain(uint32_t li) {
    Loop* l = loops_[li].get();
    epoll_event events[128];
    std::vector<int> fds;
    std::vector<Done> done;
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
            const uint32_t index = static_cast<uint32_t>(d.conn_id & 0xFFFFFFFFu);
            if (index >= l->conns.size()) continue;
            Conn& c = l->conns[index];
            if (c.fd < 0 || (c.gen & 0xFFFFu) != ((d.conn_id >> 32) & 0xFFFFu)) continue;  // closed meanwhile
            SendAll(c.fd, d.response);
            c.busy = false;
            Serve(li, l, index);
          }
          done.clear();
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
        if (closed) {
          CloseConn(l, c);
          continue;
        }
        Serve(li, l, tag);
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

#endif  // TB200_CSRC_HTTP_SERVER_H_
