// Rollout-manager core logic that needs no sockets: response merging for token-level continuation,
// sampling-budget adjustment, CIDR filtering, the local/remote load-balance controller and TOML-lite config.
//
// Behavioural parity with the reference's Rust manager: rollout-manager/src/utils.rs:45-86 (merge),
// :140-182 (extend input ids), :256-291 (shrink budget, shortest partial for grouped requests), :305 (CIDR);
// rollout-manager/src/balance.rs:23-213 (controller); rollout-manager/src/config.rs:43-123 (config order).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "json.hpp"

namespace prl {

// ------------------------------------------------------------------------------------------ responses
inline size_t count_single_response_tokens(const Json& r) {
  return r.at_path({"meta_info", "output_token_logprobs"}).size();
}
// sum=true: total tokens over sub-responses; sum=false: the shortest sub-response
inline size_t count_output_tokens(const Json& resp, bool sum) {
  if (resp.is_array()) {
    size_t acc = sum ? 0 : SIZE_MAX;
    for (auto& it : resp.arr()) {
      size_t n = count_single_response_tokens(it);
      acc = sum ? acc + n : std::min(acc, n);
    }
    return (acc == SIZE_MAX) ? 0 : acc;
  }
  return count_single_response_tokens(resp);
}

// second := first ++ second  (logprob arrays concatenated, completion_tokens summed)
inline void merge_responses(const Json& first, Json& second) {
  const Json& f_lp = first.at_path({"meta_info", "output_token_logprobs"});
  if (!second.is_object()) second = Json::object();
  Json& meta = second["meta_info"];
  if (!meta.is_object()) meta = Json::object();
  if (f_lp.is_array()) {
    Json combined = Json::array();
    for (auto& t : f_lp.arr()) combined.push_back(t);
    if (meta["output_token_logprobs"].is_array())
      for (auto& t : meta["output_token_logprobs"].arr()) combined.push_back(t);
    meta["output_token_logprobs"] = combined;
  }
  // output_ids, when present, are concatenated the same way
  const Json& f_ids = first["output_ids"];
  if (f_ids.is_array()) {
    Json combined = Json::array();
    for (auto& t : f_ids.arr()) combined.push_back(t);
    if (second["output_ids"].is_array())
      for (auto& t : second["output_ids"].arr()) combined.push_back(t);
    second["output_ids"] = combined;
  }
  int64_t total = first.at_path({"meta_info", "completion_tokens"}).as_int(0) + meta["completion_tokens"].as_int(0);
  meta["completion_tokens"] = total;
}

inline Json build_final_response(Json last, const Json* accumulated) {
  if (!accumulated || accumulated->is_null()) return last;
  if (last.is_array() && accumulated->is_array()) {
    for (size_t i = 0; i < last.size(); ++i)
      if (i < accumulated->size()) merge_responses((*accumulated)[i], last[i]);
  } else if (!last.is_array() && !accumulated->is_array()) {
    merge_responses(*accumulated, last);
  }
  return last;
}

inline void trim_response_to_used_tokens(Json& resp, size_t used) {
  auto trim_one = [&](Json& r) {
    Json* meta = r.find("meta_info");
    if (!meta) return;
    Json* lp = meta->find("output_token_logprobs");
    if (lp && lp->is_array() && lp->arr().size() > used) lp->arr().resize(used);
    Json* ids = r.find("output_ids");
    if (ids && ids->is_array() && ids->arr().size() > used) ids->arr().resize(used);
    (*meta)["completion_tokens"] = static_cast<int64_t>(std::min<size_t>(used, (*meta)["completion_tokens"].as_int(used)));
  };
  if (resp.is_array()) for (auto& r : resp.arr()) trim_one(r); else trim_one(resp);
}

// Shrink max_new_tokens by what was already generated.  For a grouped request (n samples in one body) all
// samples are truncated to the SHORTEST partial so the group stays aligned.
inline void adjust_sampling_params_for_used_tokens(Json& sampling_params, Json& prev_response) {
  if (!sampling_params["max_new_tokens"].is_number()) return;
  int64_t max_new = sampling_params["max_new_tokens"].as_int();
  size_t used = count_output_tokens(prev_response, /*sum=*/false);
  trim_response_to_used_tokens(prev_response, used);
  sampling_params["max_new_tokens"] = std::max<int64_t>(0, max_new - static_cast<int64_t>(used));
}

// Append the token ids already generated (2nd element of each logprob triple) to the prompt(s).
inline void extend_input_ids_with_response_tokens(Json& input_ids, const Json& prev_response) {
  if (!input_ids.is_array()) return;
  auto append = [](Json& dst, const Json& one) {
    const Json& lps = one.at_path({"meta_info", "output_token_logprobs"});
    for (auto& t : lps.arr())
      if (t.is_array() && t.size() > 1) dst.push_back(t[1]);
  };
  bool batch = input_ids.size() > 0 && input_ids[static_cast<size_t>(0)].is_array();
  if (batch) {
    for (size_t i = 0; i < input_ids.size(); ++i)
      if (prev_response.is_array() && i < prev_response.size()) append(input_ids[i], prev_response[i]);
  } else {
    append(input_ids, prev_response);
  }
}

// ------------------------------------------------------------------------------------------ CIDR
inline bool parse_ipv4(const std::string& s, uint32_t& out) {
  unsigned a, b, c, d;
  char tail;
  if (sscanf(s.c_str(), "%u.%u.%u.%u%c", &a, &b, &c, &d, &tail) != 4 || a > 255 || b > 255 || c > 255 || d > 255) return false;
  out = (a << 24) | (b << 16) | (c << 8) | d;
  return true;
}
inline bool ip_in_cidr(const std::string& ip, const std::string& cidr) {
  std::string net = cidr;
  int bits = 32;
  size_t slash = cidr.find('/');
  if (slash != std::string::npos) { net = cidr.substr(0, slash); bits = atoi(cidr.c_str() + slash + 1); }
  uint32_t a, n;
  if (!parse_ipv4(ip, a) || !parse_ipv4(net, n) || bits < 0 || bits > 32) return false;
  uint32_t mask = bits == 0 ? 0 : (0xFFFFFFFFu << (32 - bits));
  return (a & mask) == (n & mask);
}
// allowed = comma separated CIDRs / plain IPs
inline std::vector<std::string> filter_ips_by_config(const std::vector<std::string>& ips, const std::string& allowed) {
  std::vector<std::string> rules;
  std::stringstream ss(allowed);
  std::string item;
  while (std::getline(ss, item, ',')) {
    item.erase(0, item.find_first_not_of(" \t"));
    item.erase(item.find_last_not_of(" \t") + 1);
    if (!item.empty()) rules.push_back(item);
  }
  std::vector<std::string> out;
  for (auto& ip : ips)
    for (auto& r : rules)
      if (ip_in_cidr(ip, r)) { out.push_back(ip); break; }
  return out;
}

// ------------------------------------------------------------------------------------------ balance
// Chooses how long colocated ("local") engines may generate before they are aborted and their partial
// generations continue on remote engines, so that trainer bubbles and remote-engine bubbles even out.
class LoadBalanceState {
 public:
  static constexpr double kAlpha = 0.8;   // EMA weight when throughput stopped growing at a stable instance count
  static constexpr double kBeta = 0.2;    // EMA weight when the instance count changed

  explicit LoadBalanceState(uint64_t initial_gen_s = 150, std::map<uint64_t, uint64_t> seed = {{1, 190}, {2, 160}, {3, 105}, {4, 70}})
      : max_local_gen_s_(initial_gen_s), optimal_(std::move(seed)) {}

  uint64_t max_local_gen_s() const { std::lock_guard<std::mutex> g(mu_); return max_local_gen_s_; }
  void set_max_local_gen_s(uint64_t v) { std::lock_guard<std::mutex> g(mu_); max_local_gen_s_ = v; }
  void update_generation_stats(uint64_t total_gen_s, double resp_len_mean, uint64_t local_gen_s) {
    std::lock_guard<std::mutex> g(mu_);
    last_total_gen_s_ = total_gen_s; last_local_gen_s_ = local_gen_s; last_resp_len_mean_ = resp_len_mean;
  }
  uint64_t last_total_gen_s() const { std::lock_guard<std::mutex> g(mu_); return last_total_gen_s_; }
  uint64_t last_local_gen_s() const { std::lock_guard<std::mutex> g(mu_); return last_local_gen_s_; }
  double last_response_length_mean() const { std::lock_guard<std::mutex> g(mu_); return last_resp_len_mean_; }
  std::map<uint64_t, uint64_t> optimal_table() const { std::lock_guard<std::mutex> g(mu_); return optimal_; }

  uint64_t adjust(uint64_t total_gen_s, uint64_t step_s, uint64_t trainer_bubble_s, uint64_t num_instances, double throughput) {
    std::lock_guard<std::mutex> g(mu_);
    const uint64_t cur = max_local_gen_s_;
    auto ema_store = [&](uint64_t key, double w) {
      auto it = optimal_.find(key);
      if (it == optimal_.end()) optimal_[key] = peak_gen_s_;
      else it->second = static_cast<uint64_t>(std::llround((1.0 - w) * static_cast<double>(it->second) + w * static_cast<double>(peak_gen_s_)));
    };
    if (num_instances != last_num_instances_) {
      last_throughput_ = 0.0;                       // the next step always counts as "growing"
      ema_store(last_num_instances_, kBeta);
      auto it = optimal_.find(num_instances);
      if (it != optimal_.end()) {
        last_num_instances_ = num_instances;
        max_local_gen_s_ = peak_gen_s_ = it->second;
        return it->second;
      }
    } else {
      if (throughput <= last_throughput_) ema_store(last_num_instances_, kAlpha);   // hill-climb: commit the peak
      else peak_gen_s_ = cur;                                                       // still improving: remember
      last_throughput_ = throughput;
    }
    const uint64_t remote_bubble = step_s > total_gen_s ? step_s - total_gen_s : 0;
    uint64_t next;
    if (trainer_bubble_s < remote_bubble) {
      const uint64_t dec = (remote_bubble - trainer_bubble_s) / 3;
      next = std::max<uint64_t>(cur > dec ? cur - dec : 0, 5);
    } else {
      next = cur + (trainer_bubble_s - remote_bubble) / 3;
    }
    last_num_instances_ = num_instances;
    max_local_gen_s_ = next;
    return next;
  }

 private:
  mutable std::mutex mu_;
  uint64_t max_local_gen_s_, last_local_gen_s_ = 0, last_total_gen_s_ = 0, peak_gen_s_ = 0, last_num_instances_ = 0;
  double last_resp_len_mean_ = 0.0, last_throughput_ = 0.0;
  std::map<uint64_t, uint64_t> optimal_;
};

// ------------------------------------------------------------------------------------------ config
struct ManagerConfig {
  std::string mooncake_transfer_device_name;
  std::string mooncake_transfer_protocol = "tcp";
  std::string allowed_sender_ips = "0.0.0.0/0";
  std::vector<std::string> weight_sender_rpyc_endpoints;
  int weight_sender_rpyc_base_port = 18861;
  int num_mooncake_groups_per_sender = 1;
  int num_mooncake_engines_per_group = 1;
  int max_assigned_batches_per_stats_check = 4;
  std::string bind_addr = "0.0.0.0:5000";
  uint64_t initial_local_gen_s = 150;
  int stats_interval_ms = 1000;
  int health_interval_ms = 2000;
  int health_timeout_s = 300;
  int request_timeout_s = 3000;
  bool wait_full_local_window = false;    // reference behaviour: always sleep the whole window
};

inline std::string toml_unquote(std::string v) {
  v.erase(0, v.find_first_not_of(" \t"));
  v.erase(v.find_last_not_of(" \t\r") + 1);
  if (v.size() >= 2 && (v.front() == '"' || v.front() == '\'')) v = v.substr(1, v.size() - 2);
  return v;
}
// Flat `key = value` TOML subset (strings, ints, bools, arrays of strings) — all the manager file uses.
inline bool load_toml_config(const std::string& path, ManagerConfig& c, std::string& err) {
  std::ifstream f(path);
  if (!f) { err = "cannot open " + path; return false; }
  std::string line;
  while (std::getline(f, line)) {
    size_t hash = line.find('#');
    if (hash != std::string::npos) line.erase(hash);
    size_t eq = line.find('=');
    if (eq == std::string::npos) continue;
    std::string k = toml_unquote(line.substr(0, eq)), v = line.substr(eq + 1);
    v.erase(0, v.find_first_not_of(" \t"));
    v.erase(v.find_last_not_of(" \t\r") + 1);
    if (k == "mooncake_transfer_device_name") c.mooncake_transfer_device_name = toml_unquote(v);
    else if (k == "mooncake_transfer_protocol") c.mooncake_transfer_protocol = toml_unquote(v);
    else if (k == "allowed_sender_ips") c.allowed_sender_ips = toml_unquote(v);
    else if (k == "weight_sender_rpyc_base_port") c.weight_sender_rpyc_base_port = atoi(v.c_str());
    else if (k == "num_mooncake_groups_per_sender") c.num_mooncake_groups_per_sender = atoi(v.c_str());
    else if (k == "num_mooncake_engines_per_group") c.num_mooncake_engines_per_group = atoi(v.c_str());
    else if (k == "max_assigned_batches_per_stats_check") c.max_assigned_batches_per_stats_check = atoi(v.c_str());
    else if (k == "initial_local_gen_s") c.initial_local_gen_s = strtoull(v.c_str(), nullptr, 10);
    else if (k == "wait_full_local_window") c.wait_full_local_window = (v == "true");
    else if (k == "weight_sender_rpyc_endpoints") {
      c.weight_sender_rpyc_endpoints.clear();
      size_t a = v.find('['), b = v.rfind(']');
      if (a == std::string::npos || b == std::string::npos) { err = "bad array for " + k; return false; }
      std::stringstream ss(v.substr(a + 1, b - a - 1));
      std::string item;
      while (std::getline(ss, item, ',')) { item = toml_unquote(item); if (!item.empty()) c.weight_sender_rpyc_endpoints.push_back(item); }
    }
  }
  return true;
}

}  // namespace prl
