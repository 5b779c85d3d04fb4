// Unit tests for the manager's socket-free core (the reference ships none for its Rust manager).
#include <cassert>
#include <cstdio>

#include "core.hpp"

using namespace prl;

static Json resp(std::initializer_list<int> toks, const char* finish = "length") {
  Json r = Json::object();
  Json lps = Json::array();
  for (int t : toks) { Json e = Json::array(); e.push_back(-0.5); e.push_back(t); e.push_back(Json()); lps.push_back(e); }
  r["text"] = "";
  r["meta_info"]["output_token_logprobs"] = lps;
  r["meta_info"]["completion_tokens"] = static_cast<int64_t>(toks.size());
  r["meta_info"]["finish_reason"]["type"] = finish;
  return r;
}

static void test_json() {
  Json j = Json::parse("{\"a\":[1,2.5,\"x\\n\\u00e9\",true,null],\"b\":{\"c\":-3}}");
  assert(j["a"][static_cast<size_t>(0)].as_int() == 1);
  assert(j["a"][static_cast<size_t>(1)].as_double() == 2.5);
  assert(j["a"][static_cast<size_t>(2)].as_string() == "x\n\xc3\xa9");
  assert(j["b"]["c"].as_int() == -3);
  Json k = Json::parse(j.dump());
  assert(k.dump() == j.dump());
  Json bad;
  assert(!Json::try_parse("{\"a\":", bad));
}

static void test_merge_and_continuation() {
  Json first = resp({1, 2, 3}, "abort"), second = resp({4, 5});
  merge_responses(first, second);
  assert(count_single_response_tokens(second) == 5);
  assert(second["meta_info"]["completion_tokens"].as_int() == 5);
  assert(second["meta_info"]["output_token_logprobs"][static_cast<size_t>(3)][static_cast<size_t>(1)].as_int() == 4);
  // grouped request: shortest partial wins, budget shrinks accordingly
  Json group = Json::array();
  group.push_back(resp({1, 2, 3, 4}));
  group.push_back(resp({7, 8}));
  Json sp = Json::object();
  sp["max_new_tokens"] = 10;
  adjust_sampling_params_for_used_tokens(sp, group);
  assert(sp["max_new_tokens"].as_int() == 8);
  assert(count_single_response_tokens(group[static_cast<size_t>(0)]) == 2);
  Json ids = Json::parse("[[10,11],[10,11]]");
  extend_input_ids_with_response_tokens(ids, group);
  assert(ids.dump() == "[[10,11,1,2],[10,11,7,8]]");
  Json single = Json::parse("[5,6]");
  extend_input_ids_with_response_tokens(single, resp({9}));
  assert(single.dump() == "[5,6,9]");
  Json fin = build_final_response(Json::parse(group.dump()), &group);
  assert(count_output_tokens(fin, true) == 8);
}

static void test_cidr() {
  assert(ip_in_cidr("10.1.2.3", "10.0.0.0/8"));
  assert(!ip_in_cidr("11.1.2.3", "10.0.0.0/8"));
  assert(ip_in_cidr("192.168.1.7", "0.0.0.0/0"));
  assert(ip_in_cidr("192.168.1.7", "192.168.1.7"));
  auto f = filter_ips_by_config({"127.0.0.1", "10.2.0.4", "172.16.0.9"}, "10.0.0.0/8, 172.16.0.0/12");
  assert(f.size() == 2 && f[0] == "10.2.0.4");
}

static void test_balance() {
  LoadBalanceState b(150);
  // instance count changes 0 -> 2: table lookup
  assert(b.adjust(100, 200, 10, 2, 5.0) == 160);
  // stable count, throughput grows: peak remembered; trainer bubble (40) > remote bubble (200-180=20) -> +6
  assert(b.adjust(180, 200, 40, 2, 6.0) == 160 + (40 - 20) / 3);
  // trainer has spare time: remote bubble 100 > trainer bubble 10 -> decrease by 30
  uint64_t cur = b.max_local_gen_s();
  assert(b.adjust(100, 200, 10, 2, 7.0) == cur - 30);
  // floor at 5 s
  b.set_max_local_gen_s(8);
  assert(b.adjust(0, 1000, 0, 2, 8.0) == 5);
  // unknown instance count falls back to elastic adjustment (no table entry for 7)
  uint64_t before = b.max_local_gen_s();
  uint64_t nxt = b.adjust(50, 60, 40, 7, 1.0);
  assert(nxt == before + (40 - 10) / 3);
}

static void test_toml() {
  FILE* f = fopen("/tmp/prl_mgr_test.toml", "w");
  fputs("# comment\nmooncake_transfer_protocol = \"tcp\"\nallowed_sender_ips = \"10.0.0.0/8\"\n"
        "weight_sender_rpyc_endpoints = [\"10.0.0.1:18861\", \"10.0.0.2:18861\"]\nmax_assigned_batches_per_stats_check = 7\n", f);
  fclose(f);
  ManagerConfig c;
  std::string err;
  assert(load_toml_config("/tmp/prl_mgr_test.toml", c, err));
  assert(c.allowed_sender_ips == "10.0.0.0/8" && c.weight_sender_rpyc_endpoints.size() == 2 && c.max_assigned_batches_per_stats_check == 7);
}

int main() {
  test_json();
  test_merge_and_continuation();
  test_cidr();
  test_balance();
  test_toml();
  printf("manager core tests: OK\n");
  return 0;
}
