// polyrl_b200 rollout manager — native C++ re-implementation of the reference's Rust/tokio/axum service
// (rollout-manager/src/{main,state,handlers,instance_manager}.rs).  Same 13 REST routes and JSON schemas
// (SURVEY Appendix A), same scheduling semantics:
//   * per-stats-tick admission quota + round-robin among engines whose queue is empty, otherwise park
//   * token-level fault-tolerant continuation: on stream failure / abort keep the partial generation, evict
//     the (remote) engine, append the generated tokens to the prompt, shrink max_new_tokens, re-dispatch (<=5x)
//   * time-boxed generation on colocated ("local") engines, then abort-all and continue remotely
//   * weight-version protocol (update_weight_version / get_receive_instances CAS / update_weights activate)
//   * load-balance controller fed by /update_metrics
// Differences (documented in DESIGN.md): the local window ends early when every request already finished
// (reference always sleeps the whole budget; opt back in with wait_full_local_window=true).
#include <signal.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <ctime>
#include <deque>
#include <memory>
#include <mutex>
#include <set>
#include <shared_mutex>

#include "core.hpp"
#include "http.hpp"

using namespace prl;
using Clock = std::chrono::steady_clock;

static int g_log_level = 1;   // 0 error, 1 info, 2 debug
static void logf(int lvl, const char* fmt, ...) {
  if (lvl > g_log_level) return;
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  char ts[32];
  time_t t = time(nullptr);
  strftime(ts, sizeof ts, "%H:%M:%S", localtime(&t));
  fprintf(stderr, "[%s %s rollout-manager] ", ts, lvl == 0 ? "ERROR" : (lvl == 1 ? "INFO" : "DEBUG"));
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stderr, fmt, ap);
  va_end(ap);
  fputc('\n', stderr);
}

struct Instance {
  std::string id, host;
  int port = 0;
  bool is_local = false;
  std::string handshake_addr;          // "" for local instances
  std::string weight_sender_endpoint;  // "" for local instances
  std::string endpoint() const { return host + ":" + std::to_string(port); }
};

struct InstanceState {
  Instance inst;
  std::atomic<int64_t> running{0}, queued{0}, pending{0};
  std::atomic<uint64_t> weight_version{0};
  std::atomic<bool> updating_weight{false};
  std::atomic<bool> healthy{false};
  std::atomic<double> last_gen_throughput{0.0};
};
using InstPtr = std::shared_ptr<InstanceState>;

class AppState {
 public:
  ManagerConfig cfg;
  LoadBalanceState balance;
  std::atomic<uint64_t> latest_weight_version{0};
  std::atomic<bool> shutting_down{false};
  WorkerPool* gen_pool = nullptr;          // bounded pool for the blocking per-request engine streams (set in main)

  explicit AppState(const ManagerConfig& c) : cfg(c), balance(c.initial_local_gen_s) {}

  // ---- registry -----------------------------------------------------------------------------
  InstPtr find(const std::string& ep) {
    std::shared_lock<std::shared_mutex> g(mu_);
    auto it = instances_.find(ep);
    return it == instances_.end() ? nullptr : it->second;
  }
  bool add_instance(const Instance& i) {
    std::unique_lock<std::shared_mutex> g(mu_);
    if (instances_.count(i.endpoint())) return false;
    auto st = std::make_shared<InstanceState>();
    st->inst = i;
    instances_[i.endpoint()] = st;
    return true;
  }
  void remove_instance(const std::string& ep) {
    std::unique_lock<std::shared_mutex> g(mu_);
    instances_.erase(ep);
    active_.erase(std::remove(active_.begin(), active_.end(), ep), active_.end());
  }
  void activate(const std::string& ep) {
    {
      std::unique_lock<std::shared_mutex> g(mu_);
      if (!instances_.count(ep)) return;
      if (std::find(active_.begin(), active_.end(), ep) == active_.end()) active_.push_back(ep);
    }
    notify();
  }
  void deactivate(const std::string& ep) {
    std::unique_lock<std::shared_mutex> g(mu_);
    active_.erase(std::remove(active_.begin(), active_.end(), ep), active_.end());
  }
  void deactivate_locals() {
    std::unique_lock<std::shared_mutex> g(mu_);
    active_.erase(std::remove_if(active_.begin(), active_.end(), [&](const std::string& ep) {
      auto it = instances_.find(ep);
      return it != instances_.end() && it->second->inst.is_local;
    }), active_.end());
  }
  // new weight version: only local instances stay active (they share the trainer's weights)
  uint64_t bump_weight_version() {
    std::unique_lock<std::shared_mutex> g(mu_);
    for (auto& kv : instances_)
      if (!kv.second->inst.is_local && kv.second->pending.load() > 0)
        logf(0, "instance %s has %ld pending requests during weight version update", kv.first.c_str(), (long)kv.second->pending.load());
    active_.clear();
    uint64_t v = latest_weight_version.fetch_add(1) + 1;
    for (auto& kv : instances_)
      if (kv.second->inst.is_local) { active_.push_back(kv.first); kv.second->weight_version.store(v); }
    g.unlock();
    notify();
    return v;
  }
  std::vector<InstPtr> all() {
    std::shared_lock<std::shared_mutex> g(mu_);
    std::vector<InstPtr> out;
    for (auto& kv : instances_) out.push_back(kv.second);
    return out;
  }
  std::vector<std::string> active() {
    std::shared_lock<std::shared_mutex> g(mu_);
    return active_;
  }
  size_t num_active() { std::shared_lock<std::shared_mutex> g(mu_); return active_.size(); }

  // ---- scheduler ----------------------------------------------------------------------------
  // Admit at most max_assigned_batches_per_stats_check requests per stats tick; pick round-robin among
  // active engines with an empty queue; otherwise park until the next tick / registry change.
  InstPtr next_instance(int wait_budget_ms = -1) {
    auto deadline = Clock::now() + std::chrono::milliseconds(wait_budget_ms < 0 ? 0 : wait_budget_ms);
    for (;;) {
      if (shutting_down.load()) return nullptr;
      {
        std::shared_lock<std::shared_mutex> g(mu_);
        if (!active_.empty()) {
          uint64_t cur = assigned_.fetch_add(1);
          if (cur < static_cast<uint64_t>(cfg.max_assigned_batches_per_stats_check)) {
            std::vector<InstPtr> zero;
            for (auto& ep : active_) {
              auto it = instances_.find(ep);
              if (it != instances_.end() && it->second->queued.load() == 0) zero.push_back(it->second);
            }
            if (!zero.empty()) {
              InstPtr pick = zero[rr_.fetch_add(1) % zero.size()];
              pick->pending.fetch_add(1);
              return pick;
            }
          }
        }
      }
      std::unique_lock<std::mutex> lk(cv_mu_);
      uint64_t gen = notify_gen_;
      if (wait_budget_ms >= 0) {
        if (!cv_.wait_until(lk, deadline, [&] { return notify_gen_ != gen || shutting_down.load(); })) return nullptr;
      } else {
        cv_.wait_for(lk, std::chrono::milliseconds(2000), [&] { return notify_gen_ != gen || shutting_down.load(); });
      }
    }
  }
  void notify() {
    { std::lock_guard<std::mutex> g(cv_mu_); ++notify_gen_; }
    cv_.notify_all();
  }
  void reset_assigned() { assigned_.store(0); }

  // ---- weight senders --------------------------------------------------------------------------
  bool have_senders() { std::lock_guard<std::mutex> g(sender_mu_); return !cfg.weight_sender_rpyc_endpoints.empty(); }
  void set_senders(const std::vector<std::string>& eps, int groups, int engines) {
    {
      std::lock_guard<std::mutex> g(sender_mu_);
      cfg.weight_sender_rpyc_endpoints = eps;
      if (groups > 0) cfg.num_mooncake_groups_per_sender = groups;
      if (engines > 0) cfg.num_mooncake_engines_per_group = engines;
    }
    sender_cv_.notify_all();
  }
  bool wait_for_senders(int timeout_s) {
    std::unique_lock<std::mutex> g(sender_mu_);
    return sender_cv_.wait_for(g, std::chrono::seconds(timeout_s), [&] { return !cfg.weight_sender_rpyc_endpoints.empty() || shutting_down.load(); }) &&
           !cfg.weight_sender_rpyc_endpoints.empty();
  }
  // round-robin over (sender x group)
  void next_weight_sender(std::string& ep, int& group_idx, int& engines) {
    std::lock_guard<std::mutex> g(sender_mu_);
    size_t n = cfg.weight_sender_rpyc_endpoints.size();
    size_t total = n * static_cast<size_t>(cfg.num_mooncake_groups_per_sender);
    size_t c = sender_counter_++ % total;
    group_idx = static_cast<int>(c / n);
    ep = cfg.weight_sender_rpyc_endpoints[c % n];
    engines = cfg.num_mooncake_engines_per_group;
  }

 private:
  std::shared_mutex mu_;
  std::map<std::string, InstPtr> instances_;
  std::vector<std::string> active_;
  std::atomic<uint64_t> assigned_{0}, rr_{0};
  std::mutex cv_mu_;
  std::condition_variable cv_;
  uint64_t notify_gen_ = 0;
  std::mutex sender_mu_;
  std::condition_variable sender_cv_;
  size_t sender_counter_ = 0;
};

// ------------------------------------------------------------------------------------------------
// engine I/O
// ------------------------------------------------------------------------------------------------
struct GenError {
  int status = 502;
  std::string error, error_type;
  Json partial;      // null when nothing was produced
};

// POST {engine}/generate with stream=true, parse SSE `data:` lines.  Chunks are cumulative; keep the latest
// per sub-request index.  Returns true with the merged final response, or false with error (+partial).
static bool collect_streaming_response(AppState& st, const InstPtr& inst, Json body, const Json* accumulated, Json& out,
                                       GenError& err) {
  body["stream"] = true;
  if (accumulated && !accumulated->is_null()) extend_input_ids_with_response_tokens(body["input_ids"], *accumulated);
  const Json& ids = body["input_ids"];
  const bool is_batch = ids.is_array() && ids.size() > 0 && ids[static_cast<size_t>(0)].is_array();
  const size_t n_sub = is_batch ? ids.size() : 1;
  std::vector<Json> latest(n_sub);
  std::vector<bool> got(n_sub, false);
  bool done = false, aborted = false, decode_err = false;
  auto on_line = [&](const std::string& line) -> bool {
    if (line.rfind("data:", 0) != 0) return true;
    std::string payload = line.substr(5);
    payload.erase(0, payload.find_first_not_of(' '));
    if (payload == "[DONE]") { done = true; return false; }
    Json chunk;
    if (!Json::try_parse(payload, chunk)) { decode_err = true; return false; }
    size_t idx = is_batch ? static_cast<size_t>(chunk["index"].as_int(0)) : 0;
    if (idx >= n_sub) return true;
    if (chunk.at_path({"meta_info", "finish_reason", "type"}).as_string() == "abort") aborted = true;
    latest[idx] = std::move(chunk);
    got[idx] = true;
    return true;
  };
  HttpResult r = http_request("POST", inst->inst.host, inst->inst.port, "/generate", body.dump(),
                              st.cfg.request_timeout_s * 1000, on_line);
  auto partial_json = [&]() -> Json {
    bool any = false;
    for (bool g : got) any = any || g;
    if (!any) return accumulated ? *accumulated : Json();
    Json cur;
    if (is_batch) {
      cur = Json::array();
      for (size_t i = 0; i < n_sub; ++i) cur.push_back(got[i] ? latest[i] : Json::object());
    } else {
      cur = latest[0];
    }
    return build_final_response(cur, accumulated);
  };
  if (r.status != 0 && r.status != 200) {
    err.status = 502; err.error_type = "request_error";
    err.error = "engine returned HTTP " + std::to_string(r.status);
    return false;
  }
  if (decode_err) { err.error_type = "decode_error"; err.error = "bad SSE chunk"; err.partial = partial_json(); return false; }
  if (!done) {
    err.error_type = r.error_type.empty() ? "connection_error" : r.error_type;
    err.error = r.error.empty() ? "stream ended before [DONE]" : r.error;
    err.partial = partial_json();
    return false;
  }
  if (aborted) { err.error_type = "stream_abort"; err.error = "engine aborted the request"; err.partial = partial_json(); return false; }
  for (size_t i = 0; i < n_sub; ++i)
    if (!got[i]) { err.error_type = "decode_error"; err.error = "missing sub-response"; err.partial = partial_json(); return false; }
  out = partial_json();
  return true;
}

static void fire_and_forget_post(const std::string& host, int port, const std::string& path, const std::string& body = "") {
  std::thread([=] { http_request("POST", host, port, path, body, 5000); }).detach();
}

// One request (possibly a group of n samples) with retry / continuation.
static bool process_single_generate_request(AppState& st, const Json& original, Json& out, GenError& err) {
  Json accumulated;
  const int kMaxAttempts = 5;
  for (int attempt = 0;; ++attempt) {
    if (attempt >= kMaxAttempts) { err.status = 503; err.error = "max attempts exceeded"; err.error_type = "max_attempts"; return false; }
    InstPtr inst = st.next_instance(-1);
    if (!inst) { err.status = 503; err.error = "no rollout instances registered"; err.error_type = "no_instance"; return false; }
    Json body = original;
    if (!accumulated.is_null()) {
      adjust_sampling_params_for_used_tokens(body["sampling_params"], accumulated);
      if (body["sampling_params"]["max_new_tokens"].is_number() && body["sampling_params"]["max_new_tokens"].as_int() <= 0) {
        inst->pending.fetch_sub(1);
        out = accumulated;      // budget exhausted by the partial: it is the final answer
        return true;
      }
    }
    GenError e;
    bool ok = collect_streaming_response(st, inst, body, accumulated.is_null() ? nullptr : &accumulated, out, e);
    inst->pending.fetch_sub(1);
    if (ok) return true;
    static const std::set<std::string> retryable = {"timeout", "connection_error", "decode_error", "stream_abort", "request_error"};
    if (attempt < kMaxAttempts - 1 && retryable.count(e.error_type)) {
      logf(1, "stream failed on %s (%s); continuing elsewhere", inst->inst.endpoint().c_str(), e.error_type.c_str());
      if (!inst->inst.is_local) {     // local engines are removed by the time-box logic, never evicted here
        st.remove_instance(inst->inst.endpoint());
        fire_and_forget_post(inst->inst.host, inst->inst.port, "/shutdown?graceful=false");
      }
      if (!e.partial.is_null() && e.error_type != "request_error") accumulated = e.partial;
      continue;
    }
    err = e;
    if (!e.partial.is_null()) err.partial = e.partial;
    return false;
  }
}

static Json abort_local_requests(AppState& st) {
  Json results = Json::array();
  for (auto& i : st.all()) {
    if (!i->inst.is_local) continue;
    HttpResult r = http_request("POST", i->inst.host, i->inst.port, "/abort_request", "{\"rid\":\"\",\"abort_all\":true}", 10000);
    Json o = Json::object();
    o["endpoint"] = i->inst.endpoint();
    o["success"] = r.status == 200;
    results.push_back(o);
  }
  return results;
}

// ------------------------------------------------------------------------------------------------
// background workers
// ------------------------------------------------------------------------------------------------
static void stats_worker(AppState* st) {
  while (!st->shutting_down.load()) {
    for (auto& i : st->all()) {
      if (!i->healthy.load() && !i->inst.is_local) continue;
      HttpResult r = http_request("GET", i->inst.host, i->inst.port, "/get_server_info", "", 2000, nullptr, 1000);
      Json j;
      if (r.status == 200 && Json::try_parse(r.body, j)) {
        const Json& s = j["internal_states"][static_cast<size_t>(0)];
        i->running.store(s["#running_req"].as_int(0));
        i->queued.store(s["#queue_req"].as_int(0));
        i->last_gen_throughput.store(s["last_gen_throughput"].as_double(0.0));
      }
    }
    st->reset_assigned();
    st->notify();
    std::this_thread::sleep_for(std::chrono::milliseconds(st->cfg.stats_interval_ms));
  }
}

// Health-gate a newly registered remote engine; it only becomes *active* after a weight update.
static void health_check_instance(AppState* st, std::string ep) {
  auto deadline = Clock::now() + std::chrono::seconds(st->cfg.health_timeout_s);
  while (Clock::now() < deadline && !st->shutting_down.load()) {
    InstPtr i = st->find(ep);
    if (!i) return;
    HttpResult r = http_request("GET", i->inst.host, i->inst.port, "/health_generate", "", 5000, nullptr, 2000);
    if (r.status == 200) {
      i->healthy.store(true);
      logf(1, "instance %s is healthy", ep.c_str());
      return;
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(st->cfg.health_interval_ms));
  }
  logf(0, "instance %s failed the health gate; removing", ep.c_str());
  st->remove_instance(ep);
}

// ------------------------------------------------------------------------------------------------
// handlers
// ------------------------------------------------------------------------------------------------
static std::string gen_uuid() {
  static std::atomic<uint64_t> ctr{1};
  char buf[64];
  snprintf(buf, sizeof buf, "%08lx-%04x-4%03x-8%03x-%012lx", (unsigned long)time(nullptr), (unsigned)(ctr.load() & 0xffff),
           (unsigned)(rand() & 0xfff), (unsigned)(rand() & 0xfff), (unsigned long)ctr.fetch_add(1));
  return buf;
}
static Json ok_json() { Json o = Json::object(); o["success"] = true; return o; }
static Json err_json(const std::string& m) { Json o = Json::object(); o["success"] = false; o["error"] = m; return o; }

static void install_routes(HttpServer& srv, AppState& st) {
  srv.route("GET", "/health", [&](const HttpRequest&, HttpResponder& rsp) {
    Json o = Json::object();
    o["status"] = "healthy";
    o["message"] = "rollout manager is running";
    rsp.send(200, o.dump());
  });

  srv.route("GET", "/get_instances_status", [&](const HttpRequest&, HttpResponder& rsp) {
    Json ready = Json::array(), pending = Json::array();
    auto act = st.active();
    for (auto& i : st.all()) {
      bool is_active = std::find(act.begin(), act.end(), i->inst.endpoint()) != act.end();
      (is_active ? ready : pending).push_back(i->inst.endpoint());
    }
    Json o = Json::object();
    o["ready_instances"] = static_cast<int64_t>(ready.size());
    o["pending_instances"] = static_cast<int64_t>(pending.size());
    o["ready_instance_endpoints"] = ready;
    o["pending_instance_endpoints"] = pending;
    o["latest_weight_version"] = static_cast<int64_t>(st.latest_weight_version.load());
    o["max_local_gen_s"] = static_cast<int64_t>(st.balance.max_local_gen_s());
    rsp.send(200, o.dump());
  });

  srv.route("POST", "/register_rollout_instance", [&](const HttpRequest& req, HttpResponder& rsp) {
    Json b;
    if (!Json::try_parse(req.body, b) || !b["host"].is_string() || !b["port"].is_number()) { rsp.send(400, err_json("host and port required").dump()); return; }
    Instance i;
    i.id = gen_uuid();
    i.host = b["host"].as_string();
    i.port = static_cast<int>(b["port"].as_int());
    i.is_local = false;
    i.handshake_addr = i.host + ":" + std::to_string(b["mooncake_handshake_port"].as_int(21000));
    if (st.find(i.endpoint())) { rsp.send(409, err_json("instance already registered or pending").dump()); return; }
    if (!st.wait_for_senders(st.cfg.health_timeout_s)) { rsp.send(503, err_json("no weight senders available").dump()); return; }
    int group_idx = 0, engines = 1;
    st.next_weight_sender(i.weight_sender_endpoint, group_idx, engines);
    if (!st.add_instance(i)) { rsp.send(409, err_json("instance already registered or pending").dump()); return; }
    std::thread(health_check_instance, &st, i.endpoint()).detach();
    Json o = Json::object();
    o["mooncake_transfer_device_name"] = st.cfg.mooncake_transfer_device_name;
    o["mooncake_transfer_protocol"] = st.cfg.mooncake_transfer_protocol;
    o["weight_sender_rpyc_endpoint"] = i.weight_sender_endpoint;
    o["sender_group_idx"] = group_idx;
    o["num_mooncake_engines_per_group"] = engines;
    logf(1, "registered remote instance %s (sender %s group %d)", i.endpoint().c_str(), i.weight_sender_endpoint.c_str(), group_idx);
    rsp.send(200, o.dump());
  });

  srv.route("POST", "/register_local_rollout_instances", [&](const HttpRequest& req, HttpResponder& rsp) {
    Json b;
    if (!Json::try_parse(req.body, b) || !b.is_array()) { rsp.send(400, err_json("expected [[host, port], ...]").dump()); return; }
    for (auto& e : b.arr()) {
      if (!e.is_array() || e.size() < 2) continue;
      Instance i;
      i.id = gen_uuid();
      i.host = e[static_cast<size_t>(0)].as_string();
      i.port = static_cast<int>(e[static_cast<size_t>(1)].as_int());
      i.is_local = true;
      if (st.add_instance(i)) {
        auto p = st.find(i.endpoint());
        p->healthy.store(true);
        p->weight_version.store(st.latest_weight_version.load());
        st.activate(i.endpoint());
        logf(1, "registered local instance %s", i.endpoint().c_str());
      }
    }
    rsp.send(200, ok_json().dump());
  });

  srv.route("POST", "/generate", [&](const HttpRequest& req, HttpResponder& rsp) {
    Json b, out;
    if (!Json::try_parse(req.body, b)) { rsp.send(400, err_json("invalid json").dump()); return; }
    GenError e;
    if (process_single_generate_request(st, b, out, e)) { rsp.send(200, out.dump()); return; }
    Json o = Json::object();
    o["error"] = e.error;
    o["error_type"] = e.error_type;
    if (!e.partial.is_null()) o["partial_response"] = e.partial;
    rsp.send(e.status, o.dump());
  });

  srv.route("POST", "/batch_generate_requests", [&](const HttpRequest& req, HttpResponder& rsp) {
    Json b;
    if (!Json::try_parse(req.body, b) || !b.is_array() || b.size() == 0) { rsp.send(400, err_json("no requests provided").dump()); return; }
    struct Line {
      std::string text;
      Clock::time_point when;
      size_t tokens = 0, subs = 0;             // counted once, where the result object still exists (no re-parse of the line)
    };
    struct Shared {
      std::mutex mu;
      std::condition_variable cv;
      std::deque<Line> results;
      size_t finished = 0;
    };
    auto sh = std::make_shared<Shared>();
    const size_t total = b.size();
    const auto t_start = Clock::now();
    for (auto& item : b.arr()) {
      int64_t id = item[static_cast<size_t>(0)].as_int();
      Json body = item[static_cast<size_t>(1)];
      st.gen_pool->submit([&st, sh, id, body] {
        Json out, line = Json::object();
        GenError e;
        line["id"] = id;
        Line res;
        if (process_single_generate_request(st, body, out, e)) {
          res.tokens = count_output_tokens(out, true);
          res.subs = out.is_array() ? out.size() : 1;
          line["success"] = true;
          line["data"] = out;
        } else {
          line["success"] = false;
          line["status_code"] = e.status;
          Json eo = Json::object();
          eo["error"] = e.error;
          eo["error_type"] = e.error_type;
          if (!e.partial.is_null()) eo["partial_response"] = e.partial;
          line["error"] = eo;
        }
        res.text = line.dump() + "\n";
        res.when = Clock::now();
        { std::lock_guard<std::mutex> g(sh->mu); sh->results.push_back(std::move(res)); ++sh->finished; }
        sh->cv.notify_all();
      });
    }
    // ---- time-boxed local generation window
    const uint64_t max_gen_s = st.balance.max_local_gen_s();
    {
      std::unique_lock<std::mutex> lk(sh->mu);
      auto until = t_start + std::chrono::seconds(max_gen_s);
      if (st.cfg.wait_full_local_window) sh->cv.wait_until(lk, until, [] { return false; });
      else sh->cv.wait_until(lk, until, [&] { return sh->finished == total; });
    }
    bool all_done;
    { std::lock_guard<std::mutex> g(sh->mu); all_done = sh->finished == total; }
    if (!all_done) {
      st.deactivate_locals();
      abort_local_requests(st);
      std::this_thread::sleep_for(std::chrono::milliseconds(st.cfg.wait_full_local_window ? 1000 : 50));
    }
    const uint64_t local_gen_s = static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::seconds>(Clock::now() - t_start).count());
    // ---- stream: notifier first, then results in completion order
    rsp.begin_stream(200, "application/x-ndjson");
    rsp.write_chunk("{\"id\":0,\"success\":true}\n");
    size_t sent = 0, total_tokens = 0, total_sub = 0;
    Clock::time_point latest = t_start;
    while (sent < total) {
      Line item;
      {
        std::unique_lock<std::mutex> lk(sh->mu);
        sh->cv.wait(lk, [&] { return !sh->results.empty(); });
        item = std::move(sh->results.front());
        sh->results.pop_front();
      }
      latest = std::max(latest, item.when);
      total_tokens += item.tokens;
      total_sub += item.subs;
      if (!rsp.write_chunk(item.text)) break;    // client went away
      ++sent;
    }
    rsp.end_stream();
    if (total_sub > 0) {
      uint64_t total_gen_s = static_cast<uint64_t>(std::chrono::duration_cast<std::chrono::seconds>(latest - t_start).count());
      st.balance.update_generation_stats(total_gen_s, static_cast<double>(total_tokens) / static_cast<double>(total_sub), local_gen_s);
    }
  });

  srv.route("POST", "/update_weight_version", [&](const HttpRequest&, HttpResponder& rsp) {
    uint64_t v = st.bump_weight_version();
    logf(1, "weight version -> %lu", (unsigned long)v);
    Json o = ok_json();
    o["new_weight_version"] = static_cast<int64_t>(v);
    rsp.send(200, o.dump());
  });

  srv.route("POST", "/get_receive_instances", [&](const HttpRequest& req, HttpResponder& rsp) {
    Json b;
    if (!Json::try_parse(req.body, b)) { rsp.send(400, err_json("invalid json").dump()); return; }
    const std::string sender = b["weight_sender_endpoint"].as_string();
    const uint64_t sv = static_cast<uint64_t>(b["sender_weight_version"].as_int());
    const uint64_t latest = st.latest_weight_version.load();
    Json list = Json::array();
    int status = 200;
    if (sv > latest) {
      status = 409;
      logf(0, "weight sender version %lu is ahead of latest %lu", (unsigned long)sv, (unsigned long)latest);
    } else if (sv == latest) {
      for (auto& i : st.all()) {
        if (i->inst.is_local || i->inst.weight_sender_endpoint != sender || !i->healthy.load()) continue;
        uint64_t cur = i->weight_version.load();
        if (cur >= latest) continue;
        bool expected = false;
        if (!i->updating_weight.compare_exchange_strong(expected, true)) continue;    // someone else is updating it
        Json inst = Json::object();
        inst["id"] = i->inst.id;
        inst["endpoint"] = i->inst.endpoint();
        inst["mooncake_handshake_addr"] = i->inst.handshake_addr;
        inst["weight_sender_endpoint"] = i->inst.weight_sender_endpoint;
        Json e = Json::object();
        e["instance"] = inst;
        e["current_weight_version"] = static_cast<int64_t>(cur);
        list.push_back(e);
      }
    }
    Json o = Json::object();
    o["instances"] = list;
    rsp.send(status, o.dump());
  });

  srv.route("POST", "/update_weights", [&](const HttpRequest& req, HttpResponder& rsp) {
    Json b;
    if (!Json::try_parse(req.body, b) || !b["instance_endpoints"].is_array()) { rsp.send(400, err_json("instance_endpoints required").dump()); return; }
    const uint64_t version = static_cast<uint64_t>(b["weight_version"].as_int());
    const bool bootstrap = b["bootstrap"].as_bool(false);
    Json fwd = Json::object();
    fwd["tensors_meta"] = b["tensors_meta"];
    fwd["load_format"] = b["load_format"];
    fwd["flush_cache"] = b["flush_cache"].is_null() ? Json(true) : b["flush_cache"];
    fwd["bootstrap"] = bootstrap;
    fwd["weight_version"] = static_cast<int64_t>(version);
    const std::string payload = fwd.dump();
    Json details = Json::array();
    bool all_ok = true;
    std::mutex dmu;
    std::vector<std::thread> ths;
    for (auto& epj : b["instance_endpoints"].arr()) {
      std::string ep = epj.as_string();
      ths.emplace_back([&, ep] {
        Json d = Json::object();
        d["endpoint"] = ep;
        InstPtr i = st.find(ep);
        bool ok = false;
        if (i) {
          HttpResult r = http_request("POST", i->inst.host, i->inst.port, "/update_weights_from_agent", payload, st.cfg.request_timeout_s * 1000);
          Json jr;
          ok = r.status == 200 && Json::try_parse(r.body, jr) && jr["success"].as_bool(true);
          if (!ok) d["error"] = r.status ? r.body : r.error;
          if (ok && !bootstrap) {
            i->weight_version.store(version);
            i->updating_weight.store(false);
            if (version == st.latest_weight_version.load()) st.activate(ep);   // stale versions stay parked
          } else if (!ok) {
            i->updating_weight.store(false);
          }
        } else {
          d["error"] = "unknown instance";
        }
        d["success"] = ok;
        std::lock_guard<std::mutex> g(dmu);
        details.push_back(d);
        all_ok = all_ok && ok;
      });
    }
    for (auto& t : ths) t.join();
    Json o = Json::object();
    o["success"] = all_ok;
    o["details"] = details;
    rsp.send(all_ok ? 200 : 207, o.dump());
  });

  srv.route("PUT", "/update_weight_senders", [&](const HttpRequest& req, HttpResponder& rsp) {
    Json b;
    if (!Json::try_parse(req.body, b) || !b["weight_sender_ips"].is_array()) { rsp.send(400, err_json("weight_sender_ips required").dump()); return; }
    std::vector<std::string> eps;
    for (auto& node : b["weight_sender_ips"].arr()) {
      std::vector<std::string> ips;
      for (auto& ip : node.arr()) ips.push_back(ip.as_string());
      auto ok = filter_ips_by_config(ips, st.cfg.allowed_sender_ips);
      if (ok.empty()) { rsp.send(400, err_json("no allowed IP for a sender node").dump()); return; }
      eps.push_back(ok[0] + ":" + std::to_string(st.cfg.weight_sender_rpyc_base_port));
    }
    if (eps.empty()) { rsp.send(400, err_json("empty sender list").dump()); return; }
    st.set_senders(eps, static_cast<int>(b["num_mooncake_groups_per_sender"].as_int(0)),
                   static_cast<int>(b["num_mooncake_engines_per_group"].as_int(0)));
    rsp.send(200, ok_json().dump());
  });

  srv.route("POST", "/shutdown_instances", [&](const HttpRequest& req, HttpResponder& rsp) {
    Json b;
    if (!Json::try_parse(req.body, b) || !b["endpoints"].is_array()) { rsp.send(400, err_json("endpoints required").dump()); return; }
    const bool check = b["check_weight_update"].as_bool(false);
    Json done = Json::array();
    for (auto& epj : b["endpoints"].arr()) {
      std::string ep = epj.as_string();
      InstPtr i = st.find(ep);
      if (!i) continue;
      if (check && i->updating_weight.load()) continue;     // do not kill an engine in the middle of a transfer
      st.remove_instance(ep);
      fire_and_forget_post(i->inst.host, i->inst.port, "/shutdown?graceful=false");
      done.push_back(ep);
    }
    Json o = ok_json();
    o["message"] = "shutdown requested";
    o["shutdown_endpoints"] = done;
    rsp.send(200, o.dump());
  });

  srv.route("POST", "/update_metrics", [&](const HttpRequest& req, HttpResponder& rsp) {
    Json b;
    if (!Json::try_parse(req.body, b)) { rsp.send(400, err_json("invalid json").dump()); return; }
    const uint64_t step_s = static_cast<uint64_t>(b["step_time_s"].as_int());
    const uint64_t bubble_s = static_cast<uint64_t>(b["trainer_bubble_time_s"].as_int());
    const double thr = b["step_throughput"].as_double();
    const uint64_t total_gen = st.balance.last_total_gen_s(), local_gen = st.balance.last_local_gen_s();
    uint64_t n_remote = 0;
    for (auto& i : st.all()) if (!i->inst.is_local) ++n_remote;
    const uint64_t new_s = st.balance.adjust(total_gen, step_s, bubble_s, n_remote, thr);
    Json o = ok_json();
    o["new_max_gen_s"] = static_cast<int64_t>(new_s);
    o["new_num_rollout_instances"] = static_cast<int64_t>(n_remote);
    o["total_gen_time_s"] = static_cast<int64_t>(total_gen);
    o["local_gen_time_s"] = static_cast<int64_t>(local_gen);
    o["remote_wait_time_s"] = static_cast<int64_t>(step_s > total_gen ? step_s - total_gen : 0);
    o["response_length_mean"] = st.balance.last_response_length_mean();
    rsp.send(200, o.dump());
  });

  srv.route("POST", "/abort_local_requests", [&](const HttpRequest&, HttpResponder& rsp) {
    Json o = ok_json();
    o["message"] = "abort sent to local instances";
    o["results"] = abort_local_requests(st);
    rsp.send(200, o.dump());
  });
}

static HttpServer* g_srv = nullptr;
static AppState* g_state = nullptr;
static void on_signal(int) {
  if (g_state) { g_state->shutting_down.store(true); g_state->notify(); }
  if (g_srv) g_srv->stop();
}

int main(int argc, char** argv) {
  ManagerConfig cfg;
  std::string config_file;
  std::vector<std::string> cli_endpoints;
  int cli_max_assigned = -1;
  int http_workers = 64, max_inflight = 2048;
  std::string cli_device, cli_proto, cli_bind;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&](std::string& dst) { if (i + 1 < argc) dst = argv[++i]; };
    if (a == "--config-file") val(config_file);
    else if (a == "--bind-addr") val(cli_bind);
    else if (a == "--mooncake-transfer-device-name") val(cli_device);
    else if (a == "--mooncake-transfer-protocol") val(cli_proto);
    else if (a == "--max-assigned-batches-per-stats-check") { std::string v; val(v); cli_max_assigned = atoi(v.c_str()); }
    else if (a == "--weight-sender-rpyc-endpoints") { while (i + 1 < argc && argv[i + 1][0] != '-') cli_endpoints.push_back(argv[++i]); }
    else if (a == "--initial-local-gen-s") { std::string v; val(v); cfg.initial_local_gen_s = strtoull(v.c_str(), nullptr, 10); }
    else if (a == "--stats-interval-ms") { std::string v; val(v); cfg.stats_interval_ms = atoi(v.c_str()); }
    else if (a == "--health-interval-ms") { std::string v; val(v); cfg.health_interval_ms = atoi(v.c_str()); }
    else if (a == "--wait-full-local-window") cfg.wait_full_local_window = true;
    else if (a == "--http-workers") { std::string v; val(v); http_workers = std::max(4, atoi(v.c_str())); }
    else if (a == "--max-inflight-requests") { std::string v; val(v); max_inflight = std::max(8, atoi(v.c_str())); }
    else if (a == "--log-level") { std::string v; val(v); g_log_level = v == "debug" ? 2 : (v == "error" ? 0 : 1); }
    else if (a == "--help" || a == "-h") {
      printf("rollout-manager [--config-file F] [--bind-addr H:P] [--weight-sender-rpyc-endpoints E...] "
             "[--max-assigned-batches-per-stats-check N] [--initial-local-gen-s S] [--log-level info|debug|error]\n");
      return 0;
    }
  }
  uint64_t cli_init_gen = cfg.initial_local_gen_s;
  int cli_stats = cfg.stats_interval_ms, cli_health = cfg.health_interval_ms;
  bool cli_full = cfg.wait_full_local_window;
  if (!config_file.empty()) {                      // precedence: CLI > file > defaults
    std::string err;
    if (!load_toml_config(config_file, cfg, err)) { logf(0, "%s", err.c_str()); return 2; }
    if (cli_init_gen != ManagerConfig().initial_local_gen_s) cfg.initial_local_gen_s = cli_init_gen;
    cfg.stats_interval_ms = cli_stats; cfg.health_interval_ms = cli_health;
    cfg.wait_full_local_window = cfg.wait_full_local_window || cli_full;
  }
  if (!cli_bind.empty()) cfg.bind_addr = cli_bind;
  if (!cli_device.empty()) cfg.mooncake_transfer_device_name = cli_device;
  if (!cli_proto.empty()) cfg.mooncake_transfer_protocol = cli_proto;
  if (!cli_endpoints.empty()) cfg.weight_sender_rpyc_endpoints = cli_endpoints;
  if (cli_max_assigned > 0) cfg.max_assigned_batches_per_stats_check = cli_max_assigned;
  if (const char* lv = getenv("RUST_LOG")) g_log_level = std::string(lv) == "debug" ? 2 : (std::string(lv) == "error" ? 0 : 1);

  std::string host;
  int port = 5000;
  if (!split_host_port(cfg.bind_addr, host, port) && cfg.bind_addr.find(':') == std::string::npos) host = cfg.bind_addr;
  AppState state(cfg);
  WorkerPool gen_pool(static_cast<size_t>(max_inflight));      // blocking engine streams: bounded, queued beyond the bound
  state.gen_pool = &gen_pool;
  HttpServer srv(static_cast<size_t>(http_workers));
  install_routes(srv, state);
  if (!srv.listen(host, port)) { logf(0, "cannot bind %s", cfg.bind_addr.c_str()); return 1; }
  g_srv = &srv;
  g_state = &state;
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  signal(SIGPIPE, SIG_IGN);
  std::thread(stats_worker, &state).detach();
  logf(1, "listening on %s:%d (max_assigned_batches_per_stats_check=%d, local window %lus)", host.c_str(), srv.port(),
       cfg.max_assigned_batches_per_stats_check, (unsigned long)cfg.initial_local_gen_s);
  printf("ROLLOUT_MANAGER_PORT=%d\n", srv.port());
  fflush(stdout);
  srv.serve_forever();
  // Connection threads and the stats worker are detached and may still hold references to `srv` / `state`: leave without
  // running their destructors (found by the ThreadSanitizer build: route table freed under a live handler).
  fflush(nullptr);
  _exit(0);
}
