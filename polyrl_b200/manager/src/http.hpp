// Minimal HTTP/1.1 server and blocking client (content-length / chunked bodies, incremental line callbacks for SSE / NDJSON).
// Server: one epoll REACTOR thread owns the listening socket and every idle keep-alive connection (EPOLLONESHOT); a readable
// connection is handed to a BOUNDED worker pool for exactly one request and re-armed afterwards, so idle connections cost no
// thread and the thread count is bounded whatever the number of clients (reference: tokio/axum tasks,
// rollout-manager/src/handlers.rs:442-564).  POSIX sockets + epoll.
#pragma once
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <pthread.h>
#include <sys/epoll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace prl {

struct HttpRequest {
  std::string method, path, query, body;
  std::map<std::string, std::string> headers;   // lower-cased keys
};

inline std::string lower(std::string s) { for (auto& c : s) c = static_cast<char>(tolower(c)); return s; }

inline bool send_all(int fd, const char* p, size_t n) {
  while (n > 0) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k < 0) { if (errno == EINTR) continue; return false; }
    p += k; n -= static_cast<size_t>(k);
  }
  return true;
}
inline bool send_all(int fd, const std::string& s) { return send_all(fd, s.data(), s.size()); }

// Bounded pool of worker threads with small stacks.  Threads are created lazily up to `max_threads` and never exceed it; work
// beyond that queues.  Used for request handlers (bounded by --http-workers) and for the blocking per-request engine streams
// (bounded by --max-inflight-requests) — the round-1 manager spawned one detached thread per request instead.
class WorkerPool {
 public:
  explicit WorkerPool(size_t max_threads, size_t stack_bytes = 512 * 1024) : max_(max_threads ? max_threads : 1), stack_(stack_bytes) {}
  ~WorkerPool() { shutdown(); }
  void submit(std::function<void()> fn) {
    std::unique_lock<std::mutex> lk(mu_);
    q_.push_back(std::move(fn));
    if (idle_ == 0 && threads_ < max_) {
      ++threads_;
      lk.unlock();
      pthread_attr_t at;
      pthread_attr_init(&at);
      pthread_attr_setstacksize(&at, stack_);
      pthread_attr_setdetachstate(&at, PTHREAD_CREATE_DETACHED);
      pthread_t t;
      if (pthread_create(&t, &at, &WorkerPool::entry, this) != 0) { std::lock_guard<std::mutex> g(mu_); --threads_; }
      pthread_attr_destroy(&at);
      lk.lock();
    }
    cv_.notify_one();
  }
  size_t threads() { std::lock_guard<std::mutex> g(mu_); return threads_; }
  size_t peak_queue() { std::lock_guard<std::mutex> g(mu_); return peak_q_; }
  void shutdown() {
    std::unique_lock<std::mutex> lk(mu_);
    stop_ = true;
    cv_.notify_all();
    done_.wait_for(lk, std::chrono::seconds(2), [&] { return threads_ == 0; });
  }

 private:
  static void* entry(void* self) { static_cast<WorkerPool*>(self)->run(); return nullptr; }
  void run() {
    std::unique_lock<std::mutex> lk(mu_);
    for (;;) {
      ++idle_;
      cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
      --idle_;
      if (q_.empty()) break;                 // stop_ with nothing left to do
      if (q_.size() > peak_q_) peak_q_ = q_.size();
      auto fn = std::move(q_.front());
      q_.pop_front();
      lk.unlock();
      try { fn(); } catch (...) {}
      lk.lock();
    }
    --threads_;
    done_.notify_all();
  }
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::deque<std::function<void()>> q_;
  size_t max_, stack_, threads_ = 0, idle_ = 0, peak_q_ = 0;
  bool stop_ = false;
};

// Buffered reader over a socket with an optional per-read timeout (ms; <0 = block forever).
class SockReader {
 public:
  explicit SockReader(int fd, int timeout_ms = -1) : fd_(fd), timeout_ms_(timeout_ms) {}
  void set_timeout(int ms) { timeout_ms_ = ms; }
  // returns 1 ok, 0 EOF, -1 error, -2 timeout
  int fill() {
    if (timeout_ms_ >= 0) {
      pollfd p{fd_, POLLIN, 0};
      int r = ::poll(&p, 1, timeout_ms_);
      if (r == 0) return -2;
      if (r < 0) return errno == EINTR ? fill() : -1;
    }
    char tmp[65536];
    ssize_t k = ::recv(fd_, tmp, sizeof tmp, 0);
    if (k == 0) return 0;
    if (k < 0) return errno == EINTR ? fill() : -1;
    buf_.append(tmp, static_cast<size_t>(k));
    return 1;
  }
  // read a line terminated by \n (strips \r\n); rc as fill()
  int read_line(std::string& out) {
    for (;;) {
      size_t pos = buf_.find('\n', off_);
      if (pos != std::string::npos) {
        out.assign(buf_, off_, pos - off_);
        if (!out.empty() && out.back() == '\r') out.pop_back();
        off_ = pos + 1;
        compact();
        return 1;
      }
      int r = fill();
      if (r != 1) return r;
    }
  }
  int read_exact(size_t n, std::string& out) {
    while (buf_.size() - off_ < n) {
      int r = fill();
      if (r != 1) return r;
    }
    out.assign(buf_, off_, n);
    off_ += n;
    compact();
    return 1;
  }
  bool pending() const { return off_ < buf_.size(); }     // bytes of a pipelined request already buffered
  int read_some(std::string& out) {   // whatever is buffered, or one recv
    if (buf_.size() == off_) { int r = fill(); if (r != 1) return r; }
    out.assign(buf_, off_, std::string::npos);
    off_ = buf_.size();
    compact();
    return 1;
  }

 private:
  void compact() { if (off_ > (1u << 20) || off_ == buf_.size()) { buf_.erase(0, off_); off_ = 0; } }
  int fd_, timeout_ms_;
  std::string buf_;
  size_t off_ = 0;
};

// ------------------------------------------------------------------------------------------- server
class HttpResponder {
 public:
  explicit HttpResponder(int fd) : fd_(fd) {}
  bool send(int status, const std::string& body, const char* ctype = "application/json") {
    std::string h = "HTTP/1.1 " + std::to_string(status) + " " + reason(status) + "\r\nContent-Type: " + ctype +
                    "\r\nContent-Length: " + std::to_string(body.size()) + "\r\nConnection: keep-alive\r\n\r\n";
    sent_ = true;
    return send_all(fd_, h) && send_all(fd_, body);
  }
  bool begin_stream(int status = 200, const char* ctype = "application/x-ndjson") {
    std::string h = "HTTP/1.1 " + std::to_string(status) + " " + reason(status) + "\r\nContent-Type: " + ctype +
                    "\r\nTransfer-Encoding: chunked\r\nConnection: keep-alive\r\n\r\n";
    sent_ = streaming_ = true;
    return send_all(fd_, h);
  }
  bool write_chunk(const std::string& data) {
    if (data.empty()) return true;
    char hdr[32];
    int n = snprintf(hdr, sizeof hdr, "%zx\r\n", data.size());
    return send_all(fd_, hdr, static_cast<size_t>(n)) && send_all(fd_, data) && send_all(fd_, "\r\n", 2);
  }
  bool end_stream() { streaming_ = false; return send_all(fd_, "0\r\n\r\n", 5); }
  bool sent() const { return sent_; }
  static const char* reason(int s) {
    switch (s) {
      case 200: return "OK"; case 207: return "Multi-Status"; case 400: return "Bad Request"; case 404: return "Not Found";
      case 409: return "Conflict"; case 500: return "Internal Server Error"; case 502: return "Bad Gateway";
      case 503: return "Service Unavailable"; default: return "Status";
    }
  }

 private:
  int fd_;
  bool sent_ = false, streaming_ = false;
};

using Handler = std::function<void(const HttpRequest&, HttpResponder&)>;

class HttpServer {
 public:
  explicit HttpServer(size_t workers = 64) : pool_(workers) {}
  void route(const std::string& method, const std::string& path, Handler h) { routes_[method + " " + path] = std::move(h); }
  bool listen(const std::string& host, int port) {
    lfd_ = ::socket(AF_INET, SOCK_STREAM | SOCK_NONBLOCK, 0);
    if (lfd_ < 0) return false;
    int one = 1;
    setsockopt(lfd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_port = htons(static_cast<uint16_t>(port));
    a.sin_addr.s_addr = host.empty() || host == "0.0.0.0" ? INADDR_ANY : inet_addr(host.c_str());
    if (::bind(lfd_, reinterpret_cast<sockaddr*>(&a), sizeof a) < 0) return false;
    if (::listen(lfd_, 4096) < 0) return false;
    socklen_t len = sizeof a;
    getsockname(lfd_, reinterpret_cast<sockaddr*>(&a), &len);
    port_ = ntohs(a.sin_port);
    ep_ = ::epoll_create1(0);
    if (ep_ < 0) return false;
    epoll_event ev{};
    ev.events = EPOLLIN;
    ev.data.ptr = nullptr;                      // nullptr = the listening socket
    return ::epoll_ctl(ep_, EPOLL_CTL_ADD, lfd_, &ev) == 0;
  }
  int port() const { return port_; }
  size_t worker_threads() { return pool_.threads(); }
  size_t open_connections() { std::lock_guard<std::mutex> g(mu_); return conns_.size(); }
  // Reactor: accepts, and dispatches every connection that becomes readable to the worker pool for ONE request.
  void serve_forever() {
    std::vector<epoll_event> evs(256);
    while (!stop_.load()) {
      int n = ::epoll_wait(ep_, evs.data(), static_cast<int>(evs.size()), 200);
      for (int i = 0; i < n; ++i) {
        if (evs[i].data.ptr == nullptr) {
          for (;;) {
            int fd = ::accept4(lfd_, nullptr, nullptr, SOCK_CLOEXEC);
            if (fd < 0) break;
            int one = 1;
            setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
            auto c = std::make_shared<Conn>(fd);
            { std::lock_guard<std::mutex> g(mu_); conns_[fd] = c; }
            arm(c.get(), EPOLL_CTL_ADD);
          }
        } else {
          Conn* raw = static_cast<Conn*>(evs[i].data.ptr);
          std::shared_ptr<Conn> c;
          { std::lock_guard<std::mutex> g(mu_); auto it = conns_.find(raw->fd); if (it != conns_.end() && it->second.get() == raw) c = it->second; }
          if (c) pool_.submit([this, c] { this->serve_ready(c); });
        }
      }
    }
    // lfd_ / ep_ stay open: workers may still be finishing a request (epoll_ctl on ep_) while the process winds down —
    // the ThreadSanitizer build flags closing them here as a race; main() leaves through _exit()
  }
  void stop() { stop_.store(true); }

 private:
  struct Conn {
    explicit Conn(int f) : fd(f), rd(f, 300000) {}
    int fd;
    SockReader rd;
  };
  void arm(Conn* c, int op) {
    epoll_event ev{};
    ev.events = EPOLLIN | EPOLLRDHUP | EPOLLONESHOT;        // one-shot: exactly one worker owns the connection at a time
    ev.data.ptr = c;
    ::epoll_ctl(ep_, op, c->fd, &ev);
  }
  void drop(const std::shared_ptr<Conn>& c) {
    ::epoll_ctl(ep_, EPOLL_CTL_DEL, c->fd, nullptr);
    { std::lock_guard<std::mutex> g(mu_); conns_.erase(c->fd); }
    ::close(c->fd);
  }
  void serve_ready(const std::shared_ptr<Conn>& c) {
    for (;;) {
      const int rc = handle_one(*c);
      if (rc != 1) { drop(c); return; }
      if (!c->rd.pending()) break;              // pipelined request already buffered: serve it now
    }
    arm(c.get(), EPOLL_CTL_MOD);                // idle again: back to the reactor
  }
  // returns 1 = request served, keep the connection; 0 = close
  int handle_one(Conn& c) {
    SockReader& rd = c.rd;
    const int fd = c.fd;
    HttpRequest req;
    std::string line;
    do {
      if (rd.read_line(line) != 1) return 0;
    } while (line.empty());
    size_t s1 = line.find(' '), s2 = line.rfind(' ');
    if (s1 == std::string::npos || s2 == s1) return 0;
    req.method = line.substr(0, s1);
    std::string target = line.substr(s1 + 1, s2 - s1 - 1);
    size_t q = target.find('?');
    req.path = q == std::string::npos ? target : target.substr(0, q);
    req.query = q == std::string::npos ? "" : target.substr(q + 1);
    for (;;) {
      if (rd.read_line(line) != 1) return 0;
      if (line.empty()) break;
      size_t cpos = line.find(':');
      if (cpos == std::string::npos) continue;
      std::string k = lower(line.substr(0, cpos));
      size_t v = cpos + 1;
      while (v < line.size() && line[v] == ' ') ++v;
      req.headers[k] = line.substr(v);
    }
    auto te = req.headers.find("transfer-encoding");
    if (te != req.headers.end() && lower(te->second).find("chunked") != std::string::npos) {
      for (;;) {
        if (rd.read_line(line) != 1) return 0;
        size_t n = strtoul(line.c_str(), nullptr, 16);
        if (n == 0) { rd.read_line(line); break; }
        std::string part;
        if (rd.read_exact(n, part) != 1) return 0;
        req.body += part;
        rd.read_line(line);
      }
    } else {
      auto cl = req.headers.find("content-length");
      size_t n = cl == req.headers.end() ? 0 : strtoul(cl->second.c_str(), nullptr, 10);
      if (n > 0 && rd.read_exact(n, req.body) != 1) return 0;
    }
    HttpResponder resp(fd);
    auto it = routes_.find(req.method + " " + req.path);
    if (it == routes_.end()) {
      resp.send(404, "{\"error\":\"not found\"}");
    } else {
      try {
        it->second(req, resp);
      } catch (const std::exception& e) {
        if (!resp.sent()) resp.send(500, std::string("{\"error\":\"") + e.what() + "\"}");
      }
      if (!resp.sent()) resp.send(500, "{\"error\":\"handler produced no response\"}");
    }
    auto conn = req.headers.find("connection");
    if (conn != req.headers.end() && lower(conn->second) == "close") return 0;
    return 1;
  }
  std::map<std::string, Handler> routes_;
  WorkerPool pool_;
  std::mutex mu_;
  std::map<int, std::shared_ptr<Conn>> conns_;
  int lfd_ = -1, ep_ = -1, port_ = 0;
  std::atomic<bool> stop_{false};
};

// ------------------------------------------------------------------------------------------- client
struct HttpResult {
  int status = 0;               // 0 => transport failure
  std::string body;
  std::string error_type;       // "", "connection_error", "timeout", "decode_error"
  std::string error;
};

inline int connect_to(const std::string& host, int port, int timeout_ms) {
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  if (getaddrinfo(host.c_str(), std::to_string(port).c_str(), &hints, &res) != 0 || !res) return -1;
  int fd = ::socket(res->ai_family, res->ai_socktype, res->ai_protocol);
  if (fd < 0) { freeaddrinfo(res); return -1; }
  timeval tv{timeout_ms / 1000, (timeout_ms % 1000) * 1000};
  setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
  int rc = ::connect(fd, res->ai_addr, res->ai_addrlen);
  freeaddrinfo(res);
  if (rc < 0) { ::close(fd); return -1; }
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
  return fd;
}

// One-shot request (Connection: close).  If on_line is given the body is delivered line by line as it arrives
// (SSE / NDJSON) and `body` stays empty; on_line returning false stops reading early.
inline HttpResult http_request(const std::string& method, const std::string& host, int port, const std::string& path,
                               const std::string& body = "", int timeout_ms = 30000,
                               const std::function<bool(const std::string&)>& on_line = nullptr, int connect_timeout_ms = 5000) {
  HttpResult r;
  int fd = connect_to(host, port, connect_timeout_ms);
  if (fd < 0) { r.error_type = "connection_error"; r.error = "connect failed"; return r; }
  std::string req = method + " " + path + " HTTP/1.1\r\nHost: " + host + ":" + std::to_string(port) +
                    "\r\nContent-Type: application/json\r\nAccept: */*\r\nConnection: close\r\nContent-Length: " +
                    std::to_string(body.size()) + "\r\n\r\n";
  if (!send_all(fd, req) || !send_all(fd, body)) {
    ::close(fd); r.error_type = "connection_error"; r.error = "send failed"; return r;
  }
  SockReader rd(fd, timeout_ms);
  auto fail = [&](int rc, const char* what) {
    r.error_type = rc == -2 ? "timeout" : "connection_error";
    r.error = what;
    ::close(fd);
    return r;
  };
  std::string line;
  int rc = rd.read_line(line);
  if (rc != 1) return fail(rc, "no status line");
  if (line.size() < 12) return fail(-1, "bad status line");
  r.status = atoi(line.c_str() + 9);
  bool chunked = false;
  long clen = -1;
  for (;;) {
    rc = rd.read_line(line);
    if (rc != 1) { r.status = 0; return fail(rc, "headers truncated"); }
    if (line.empty()) break;
    std::string l = lower(line);
    if (l.rfind("transfer-encoding:", 0) == 0 && l.find("chunked") != std::string::npos) chunked = true;
    if (l.rfind("content-length:", 0) == 0) clen = atol(l.c_str() + 15);
  }
  std::string pending;     // partial line carried across chunks
  bool stop = false;
  auto feed = [&](const std::string& data) {
    if (!on_line) { r.body += data; return; }
    pending += data;
    size_t pos;
    while (!stop && (pos = pending.find('\n')) != std::string::npos) {
      std::string ln = pending.substr(0, pos);
      pending.erase(0, pos + 1);
      if (!ln.empty() && ln.back() == '\r') ln.pop_back();
      if (!on_line(ln)) stop = true;
    }
  };
  if (chunked) {
    for (;;) {
      rc = rd.read_line(line);
      if (rc != 1) { int st = r.status; r = fail(rc, "chunk header truncated"); r.status = st; r.status = 0; return r; }
      size_t n = strtoul(line.c_str(), nullptr, 16);
      if (n == 0) break;
      std::string part;
      rc = rd.read_exact(n, part);
      if (rc != 1) { r = fail(rc, "chunk truncated"); r.status = 0; return r; }
      feed(part);
      rd.read_line(line);
      if (stop) break;
    }
  } else if (clen >= 0) {
    std::string part;
    rc = rd.read_exact(static_cast<size_t>(clen), part);
    if (rc != 1) { r = fail(rc, "body truncated"); r.status = 0; return r; }
    feed(part);
  } else {
    for (;;) {
      std::string part;
      rc = rd.read_some(part);
      if (rc == 0) break;
      if (rc != 1) { r = fail(rc, "body read failed"); r.status = 0; return r; }
      feed(part);
      if (stop) break;
    }
  }
  if (on_line && !pending.empty() && !stop) on_line(pending);
  ::close(fd);
  return r;
}

inline bool split_host_port(const std::string& hp, std::string& host, int& port) {
  size_t c = hp.rfind(':');
  if (c == std::string::npos) return false;
  host = hp.substr(0, c);
  port = atoi(hp.c_str() + c + 1);
  return port > 0;
}

}  // namespace prl
