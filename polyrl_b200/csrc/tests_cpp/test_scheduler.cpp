// Stand-alone stress test of the native engine scheduler (csrc/scheduler.hpp), meant to be built with ASAN + UBSAN
// (tests/test_engine_server.py::test_scheduler_cpp_under_sanitizers): random add / abort / admit / prefill-plan / slot
// compaction / finish traffic with n-sample groups and the prefix cache; after every operation the slot and page
// accounting must balance.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>

#include "../scheduler.hpp"

#define CHECK(c)                                                              \
  do {                                                                        \
    if (!(c)) {                                                               \
      std::fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #c);       \
      std::exit(1);                                                           \
    }                                                                         \
  } while (0)

int main() {
  const int64_t PAGE = 16, NB = 96, SLOTS = 12;
  std::mt19937_64 rng(1234);
  auto rnd = [&](int64_t n) { return static_cast<int64_t>(rng() % static_cast<uint64_t>(n)); };
  prl::KvPageManager kv(NB, PAGE);
  prl::EngineScheduler sched(&kv, SLOTS, PAGE, true);
  std::vector<std::vector<int64_t>> prefixes;
  for (int i = 0; i < 5; ++i) {
    std::vector<int64_t> p(3 * PAGE);
    for (auto& t : p) t = rnd(50);
    prefixes.push_back(p);
  }
  struct Live { int64_t h; std::vector<int64_t> prompt; };
  std::vector<Live> live;
  std::vector<int64_t> waiting;
  int64_t admitted_total = 0, group = 0, cached_total = 0, followers = 0;
  for (int step = 0; step < 30000; ++step) {
    const int64_t op = rnd(100);
    if (op < 35) {                                            // a prompt, sometimes an n-sample group of it
      std::vector<int64_t> p = prefixes[rnd(5)];
      p.resize(rnd(4) * PAGE);
      const int64_t extra = 1 + rnd(2 * PAGE);
      for (int64_t i = 0; i < extra; ++i) p.push_back(rnd(50));
      const int64_t n = rnd(4) == 0 ? 1 + rnd(4) : 1;
      const int64_t g = n > 1 ? group++ : -1;
      for (int64_t i = 0; i < n; ++i) waiting.push_back(sched.add(p, 1 + rnd(40), g, rnd(10) != 0));
    } else if (op < 40 && !waiting.empty()) {                 // abort a waiting request
      const size_t i = static_cast<size_t>(rnd(static_cast<int64_t>(waiting.size())));
      CHECK(sched.remove_waiting(waiting[i]));
      CHECK(!sched.remove_waiting(waiting[i]));
      waiting.erase(waiting.begin() + static_cast<long>(i));
    } else if (op < 65) {                                     // admission round + the prefill plan of what entered
      const std::vector<int64_t> hs = sched.admit(64 + rnd(256));
      std::vector<std::tuple<int64_t, int64_t, int64_t>> spans;
      bool gather = false;
      for (int64_t h : hs) {
        CHECK(!waiting.empty() && waiting.front() == h);      // FIFO
        waiting.erase(waiting.begin());
        const prl::SchedRequest& r = sched.get(h);
        CHECK(r.live && r.slot >= 0 && r.slot < SLOTS);
        const int64_t plen = static_cast<int64_t>(r.prompt.size());
        CHECK(static_cast<int64_t>(r.blocks.size()) == (plen + r.max_new + PAGE - 1) / PAGE);
        CHECK(r.cached_tokens % PAGE == 0 && r.cached_tokens < plen);
        if (r.donor >= 0) {
          ++followers;
          const prl::SchedRequest& d = sched.get(r.donor);
          CHECK(d.prompt == r.prompt);
          for (int64_t i = 0; i < r.shared_prefix_blocks; ++i) CHECK(d.blocks[i] == r.blocks[i] && kv.ref(r.blocks[i]) >= 2);
        } else {
          spans.emplace_back(h, r.cached_tokens, plen);
          gather = gather || r.cached_tokens > 0;
          cached_total += r.cached_tokens;
        }
        live.push_back({h, r.prompt});
        ++admitted_total;
      }
      if (!spans.empty()) {
        const prl::PrefillArrays a = sched.prefill_arrays(spans, gather);
        int64_t total = 0;
        for (auto& s : spans) total += std::get<2>(s) - std::get<1>(s);
        CHECK(static_cast<int64_t>(a.ids.size()) == total && a.pos.size() == a.ids.size() && a.slots.size() == a.ids.size());
        CHECK(a.lens.size() == spans.size() && (gather ? a.klens.size() == spans.size() : a.g_blk.empty()));
        std::set<int64_t> seen(a.slots.begin(), a.slots.end());
        CHECK(seen.size() == a.slots.size());                 // no two new tokens share a KV slot
        for (int64_t s : a.slots) CHECK(s >= PAGE && s < NB * PAGE);   // never the scratch page
      }
    } else if (op < 70 && !live.empty()) {                    // compaction: move one request to a free slot
      const std::vector<int64_t> fs = sched.free_slots();
      if (!fs.empty()) {
        const Live& l = live[static_cast<size_t>(rnd(static_cast<int64_t>(live.size())))];
        const int64_t old_slot = sched.get(l.h).slot, dst = fs[static_cast<size_t>(rnd(static_cast<int64_t>(fs.size())))];
        sched.move_slots({{l.h, dst}});
        CHECK(sched.get(l.h).slot == dst);
        const std::vector<int64_t> after = sched.free_slots();
        CHECK(after.size() == fs.size() && std::find(after.begin(), after.end(), old_slot) != after.end());
        CHECK(std::is_sorted(after.begin(), after.end(), [](int64_t a, int64_t b) { return a > b; }));
      }
    } else if (!live.empty()) {                               // finish: publish full pages, release
      const size_t i = static_cast<size_t>(rnd(static_cast<int64_t>(live.size())));
      std::vector<int64_t> toks = live[i].prompt;
      const int64_t gen = 1 + rnd(sched.get(live[i].h).max_new);
      for (int64_t k = 0; k < gen; ++k) toks.push_back(rnd(50));
      sched.finish(live[i].h, toks, rnd(10) != 0);
      live.erase(live.begin() + static_cast<long>(i));
    }
    if (rnd(500) == 0) kv.flush();
    // invariants: every slot is either free or held by exactly one live request; pages balance
    CHECK(sched.num_free_slots() + static_cast<int64_t>(live.size()) == SLOTS);
    CHECK(sched.num_live() == static_cast<int64_t>(live.size()));
    CHECK(sched.num_waiting() == static_cast<int64_t>(waiting.size()));
    std::set<int64_t> slots;
    for (auto& l : live) CHECK(slots.insert(sched.get(l.h).slot).second);
    std::set<int64_t> used;
    for (auto& l : live)
      for (int64_t b : sched.get(l.h).blocks) { CHECK(b > 0 && kv.ref(b) >= 1); used.insert(b); }
    CHECK(kv.available() + static_cast<int64_t>(used.size()) <= NB - 1);          // the rest is held by the cache only
    CHECK(kv.available() + static_cast<int64_t>(used.size()) + kv.cached_pages() >= NB - 1);
  }
  for (auto& l : live) sched.finish(l.h, l.prompt, false);
  for (int64_t h : waiting) CHECK(sched.remove_waiting(h));
  kv.flush();
  CHECK(kv.available() == NB - 1 && sched.num_free_slots() == SLOTS && sched.num_live() == 0);
  CHECK(admitted_total > 1000 && cached_total > 0 && followers > 0);
  std::printf("SCHEDULER_OK admitted=%lld cached_tokens=%lld followers=%lld\n", static_cast<long long>(admitted_total),
              static_cast<long long>(cached_total), static_cast<long long>(followers));
  return 0;
}
