// Stand-alone stress test of the native KV page manager, meant to be built with -fsanitize=address,undefined
// (tests/test_engine_server.py::test_kv_manager_cpp_under_sanitizers): random admit / finish / evict / flush traffic with
// invariant checks after every operation.
#include <cassert>
#include <cstdio>
#include <random>
#include <vector>

#include "../kv_manager.hpp"

using prl::KvPageManager;

struct Live {
  std::vector<int64_t> toks, blocks;
};

int main() {
  const int64_t NB = 64, PAGE = 16;
  KvPageManager m(NB, PAGE);
  std::mt19937_64 rng(7);
  std::vector<Live> live;
  std::vector<std::vector<int64_t>> prefixes(4);
  for (auto& p : prefixes)
    for (int i = 0; i < 4 * PAGE; ++i) p.push_back(static_cast<int64_t>(rng() % 11));
  int64_t admitted = 0, refused = 0;
  for (int step = 0; step < 20000; ++step) {
    const double op = (rng() % 1000) / 1000.0;
    if (op < 0.5) {
      Live r;
      const auto& base = prefixes[rng() % prefixes.size()];
      const int64_t keep = static_cast<int64_t>(rng() % 5) * PAGE;
      r.toks.assign(base.begin(), base.begin() + keep);
      const int64_t tail = PAGE * (1 + rng() % 2) + rng() % PAGE;
      for (int64_t i = 0; i < tail; ++i) r.toks.push_back(static_cast<int64_t>(rng() % 11));
      const int64_t n_pages = (static_cast<int64_t>(r.toks.size()) - 1) / PAGE;
      auto keys = m.page_keys(r.toks, n_pages);
      auto hit = m.match(keys);
      m.share(hit);
      const int64_t need = (static_cast<int64_t>(r.toks.size()) + PAGE - 1) / PAGE - static_cast<int64_t>(hit.size());
      if (need > m.available()) m.evict(need - m.available());
      if (need > m.available()) {
        m.release(hit);
        ++refused;
        continue;
      }
      r.blocks = hit;
      auto fresh = m.alloc(need);
      r.blocks.insert(r.blocks.end(), fresh.begin(), fresh.end());
      live.push_back(std::move(r));
      ++admitted;
    } else if (op < 0.97) {
      if (live.empty()) continue;
      const size_t i = rng() % live.size();
      Live r = std::move(live[i]);
      live.erase(live.begin() + static_cast<long>(i));
      const int64_t n_full = (static_cast<int64_t>(r.toks.size()) - 1) / PAGE;
      if (rng() % 10 != 0) {
        std::vector<int64_t> head(r.blocks.begin(), r.blocks.begin() + n_full);
        m.insert(m.page_keys(r.toks, n_full), head);
      }
      m.release(r.blocks);
    } else {
      m.flush();
    }
    // invariants: every page is free, or referenced by a live request and / or the cache
    int64_t referenced = 0;
    for (int64_t b = 1; b < NB; ++b) {
      assert(m.ref(b) >= 0);
      referenced += m.ref(b) > 0;
    }
    assert(referenced + m.available() == NB - 1);
    assert(m.cached_pages() <= NB - 1);
  }
  for (auto& r : live) m.release(r.blocks);
  m.flush();
  assert(m.available() == NB - 1);
  std::printf("KV_MANAGER_OK admitted=%lld refused=%lld hits=%lld misses=%lld\n", static_cast<long long>(admitted),
              static_cast<long long>(refused), static_cast<long long>(m.hits()), static_cast<long long>(m.misses()));
  return 0;
}
