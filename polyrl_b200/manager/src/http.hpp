// Minimal HTTP/1.1 server (thread per connection, keep-alive, chunked streaming responses) and blocking
// client (content-length / chunked bodies, incremental line callbacks for SSE / NDJSON).  POSIX sockets.
#pragma once
#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstring>
#include <functional>
#include <map>
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
  void route(const std::string& method, const std::string& path, Handler h) { routes_[method + " " + path] = std::move(h); }
  bool listen(const std::string& host, int port) {
    lfd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    if (lfd_ < 0) return false;
    int one = 1;
    setsockopt(lfd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_port = htons(static_cast<uint16_t>(port));
    a.sin_addr.s_addr = host.empty() || host == "0.0.0.0" ? INADDR_ANY : inet_addr(host.c_str());
    if (::bind(lfd_, reinterpret_cast<sockaddr*>(&a), sizeof a) < 0) return false;
    if (::listen(lfd_, 1024) < 0) return false;
    socklen_t len = sizeof a;
    getsockname(lfd_, reinterpret_cast<sockaddr*>(&a), &len);
    port_ = ntohs(a.sin_port);
    return true;
  }
  int port() const { return port_; }
  void serve_forever() {
    while (!stop_.load()) {
      pollfd p{lfd_, POLLIN, 0};
      int r = ::poll(&p, 1, 200);
      if (r <= 0) continue;
      int fd = ::accept(lfd_, nullptr, nullptr);
      if (fd < 0) continue;
      int one = 1;
      setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
      std::thread([this, fd] { this->handle_conn(fd); }).detach();
    }
    ::close(lfd_);
  }
  void stop() { stop_.store(true); }

 private:
  void handle_conn(int fd) {
    SockReader rd(fd, 300000);
    for (;;) {
      HttpRequest req;
      std::string line;
      if (rd.read_line(line) != 1) break;
      if (line.empty()) continue;
      size_t s1 = line.find(' '), s2 = line.rfind(' ');
      if (s1 == std::string::npos || s2 == s1) break;
      req.method = line.substr(0, s1);
      std::string target = line.substr(s1 + 1, s2 - s1 - 1);
      size_t q = target.find('?');
      req.path = q == std::string::npos ? target : target.substr(0, q);
      req.query = q == std::string::npos ? "" : target.substr(q + 1);
      bool ok = true;
      for (;;) {
        if (rd.read_line(line) != 1) { ok = false; break; }
        if (line.empty()) break;
        size_t c = line.find(':');
        if (c == std::string::npos) continue;
        std::string k = lower(line.substr(0, c));
        size_t v = c + 1;
        while (v < line.size() && line[v] == ' ') ++v;
        req.headers[k] = line.substr(v);
      }
      if (!ok) break;
      auto te = req.headers.find("transfer-encoding");
      if (te != req.headers.end() && lower(te->second).find("chunked") != std::string::npos) {
        for (;;) {
          if (rd.read_line(line) != 1) { ok = false; break; }
          size_t n = strtoul(line.c_str(), nullptr, 16);
          if (n == 0) { rd.read_line(line); break; }
          std::string part;
          if (rd.read_exact(n, part) != 1) { ok = false; break; }
          req.body += part;
          rd.read_line(line);
        }
      } else {
        auto cl = req.headers.find("content-length");
        size_t n = cl == req.headers.end() ? 0 : strtoul(cl->second.c_str(), nullptr, 10);
        if (n > 0 && rd.read_exact(n, req.body) != 1) ok = false;
      }
      if (!ok) break;
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
      if (conn != req.headers.end() && lower(conn->second) == "close") break;
    }
    ::close(fd);
  }
  std::map<std::string, Handler> routes_;
  int lfd_ = -1, port_ = 0;
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
