// polyrl_b200 — native admission scheduler of the rollout engine (continuous batching over the paged KV pool).
//
// Host-only C++.  Owns, per request: the prompt, its batch slot and its KV page list; decides which waiting requests
// enter the running batch (free slot + every page the request can ever need reserved up front, so a running sequence is
// never preempted), applies the cross-request prefix cache and the n-sample prompt sharing (the first member of a group
// admitted in a round prefills; the others share its full prompt pages), and builds the flat per-token arrays of a
// prefill chunk (token ids, positions, KV slot of every new token, and — behind a cached prefix — the page / offset of
// every key position) in one pass instead of Python list comprehensions over every token of every prompt.
//
// Same decisions, in the same order, as the Python reference kept in rollout/engine.py (`RolloutEngine._admit`,
// the planning half of `_prefill`, `_finish`): tests/test_engine_server.py drives both with one random workload.
// Role in the reference stack: SGLang's scheduler (waiting queue -> running batch, radix-cache match, chunked prefill)
// behind 3rdparty/verl/verl/workers/rollout/sglang_rollout/sglang_rollout.py.
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <stdexcept>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "kv_manager.hpp"

namespace prl {

struct SchedRequest {
  std::vector<int64_t> prompt;
  int64_t max_new = 0;
  int64_t group = -1;          // n-sample group id (-1: none)
  bool cacheable = true;       // image prompts: placeholder ids do not identify the content
  int64_t slot = -1;
  std::vector<int64_t> blocks;
  int64_t shared_prefix_blocks = 0;
  int64_t cached_tokens = 0;
  int64_t donor = -1;          // handle of the group member whose prefill this request re-uses
  bool live = false;           // holds a slot / pages
};

struct PrefillArrays {
  std::vector<int64_t> ids, pos, slots, lens, g_blk, g_off, klens;
};

class EngineScheduler {
 public:
  EngineScheduler(KvPageManager* kv, int64_t max_num_seqs, int64_t page, bool prefix_cache)
      : kv_(kv), page_(page), prefix_cache_(prefix_cache) {
    if (kv == nullptr || max_num_seqs < 1 || page < 1) throw std::invalid_argument("EngineScheduler: bad arguments");
    for (int64_t s = max_num_seqs - 1; s >= 0; --s) free_slots_.push_back(s);      // popped from the back: 0, 1, 2, ...
  }

  int64_t add(std::vector<int64_t> prompt, int64_t max_new, int64_t group, bool cacheable) {
    const int64_t h = next_++;
    SchedRequest r;
    r.prompt = std::move(prompt);
    r.max_new = max_new;
    r.group = group;
    r.cacheable = cacheable;
    reqs_.emplace(h, std::move(r));
    waiting_.push_back(h);
    return h;
  }

  // drop a request that is still waiting (abort); false when it is not in the queue
  bool remove_waiting(int64_t h) {
    for (auto it = waiting_.begin(); it != waiting_.end(); ++it) {
      if (*it == h) {
        waiting_.erase(it);
        reqs_.erase(h);
        return true;
      }
    }
    return false;
  }

  int64_t num_waiting() const { return static_cast<int64_t>(waiting_.size()); }
  int64_t num_free_slots() const { return static_cast<int64_t>(free_slots_.size()); }
  int64_t num_live() const {
    int64_t n = 0;
    for (auto& kv : reqs_) n += kv.second.live ? 1 : 0;
    return n;
  }

  // Pop waiting requests that fit: a slot, every KV page reserved up front, this round's prefill token budget.
  // Returns the admitted handles in order; per-request results are read through the accessors below.
  std::vector<int64_t> admit(int64_t max_prefill_tokens) {
    std::vector<int64_t> admitted;
    std::unordered_map<int64_t, int64_t> shared;          // group -> donor admitted in THIS round
    int64_t tokens = 0;
    while (!waiting_.empty() && !free_slots_.empty()) {
      const int64_t h = waiting_.front();
      SchedRequest& r = reqs_.at(h);
      const int64_t plen = static_cast<int64_t>(r.prompt.size());
      const int64_t need = (plen + r.max_new + page_ - 1) / page_;
      const int64_t full_prompt_pages = plen / page_;
      int64_t donor = -1;
      if (r.group >= 0) {
        auto it = shared.find(r.group);
        if (it != shared.end()) donor = it->second;
      }
      std::vector<int64_t> cached;
      if (donor < 0 && prefix_cache_ && r.cacheable) {
        // at least the last prompt token is recomputed (its logits seed the first sampled token)
        cached = kv_->match(kv_->page_keys(r.prompt, (plen - 1) / page_));
        kv_->share(cached);                                // pin before any eviction below can touch them
      }
      const int64_t fresh = need - (donor >= 0 ? full_prompt_pages : static_cast<int64_t>(cached.size()));
      if (fresh > kv_->available() && prefix_cache_) kv_->evict(fresh - kv_->available());
      if (fresh > kv_->available() || (tokens >= max_prefill_tokens && !admitted.empty())) {
        if (!cached.empty()) kv_->release(cached);
        break;
      }
      waiting_.pop_front();
      if (donor >= 0) {
        const SchedRequest& d = reqs_.at(donor);
        std::vector<int64_t> pre(d.blocks.begin(), d.blocks.begin() + full_prompt_pages);
        kv_->share(pre);
        r.blocks = std::move(pre);
        r.shared_prefix_blocks = full_prompt_pages;
        r.donor = donor;
        r.cached_tokens = 0;
      } else {
        r.cached_tokens = static_cast<int64_t>(cached.size()) * page_;
        r.blocks = std::move(cached);
        if (r.group >= 0) shared[r.group] = h;
      }
      const std::vector<int64_t> got = kv_->alloc(fresh);
      r.blocks.insert(r.blocks.end(), got.begin(), got.end());
      r.slot = free_slots_.back();
      free_slots_.pop_back();
      r.live = true;
      admitted.push_back(h);
      tokens += donor >= 0 ? 0 : plen - r.cached_tokens;
    }
    return admitted;
  }

  const SchedRequest& get(int64_t h) const { return reqs_.at(h); }

  // Flat arrays of one prefill chunk: request spans[i] = (handle, c, e) contributes prompt tokens [c, e).
  // With `gather` the page / offset of EVERY key position [0, e) is emitted too (attention behind a cached prefix).
  PrefillArrays prefill_arrays(const std::vector<std::tuple<int64_t, int64_t, int64_t>>& spans, bool gather) const {
    PrefillArrays out;
    int64_t total = 0, ktotal = 0;
    for (auto& s : spans) {
      total += std::get<2>(s) - std::get<1>(s);
      ktotal += std::get<2>(s);
    }
    out.ids.reserve(total);
    out.pos.reserve(total);
    out.slots.reserve(total);
    if (gather) {
      out.g_blk.reserve(ktotal);
      out.g_off.reserve(ktotal);
    }
    for (auto& s : spans) {
      const SchedRequest& r = reqs_.at(std::get<0>(s));
      const int64_t c = std::get<1>(s), e = std::get<2>(s);
      if (c < 0 || e > static_cast<int64_t>(r.prompt.size()) || c > e) throw std::out_of_range("prefill span outside the prompt");
      if ((e + page_ - 1) / page_ > static_cast<int64_t>(r.blocks.size())) throw std::out_of_range("prefill span outside the request's pages");
      for (int64_t t = c; t < e; ++t) {
        out.ids.push_back(r.prompt[t]);
        out.pos.push_back(t);
        out.slots.push_back(r.blocks[t / page_] * page_ + t % page_);
      }
      out.lens.push_back(e - c);
      if (gather) {
        for (int64_t t = 0; t < e; ++t) {
          out.g_blk.push_back(r.blocks[t / page_]);
          out.g_off.push_back(t % page_);
        }
        out.klens.push_back(e);
      }
    }
    return out;
  }

  // Block-table rows (zero padded to `width`) of the given requests, row-major.
  std::vector<int32_t> block_rows(const std::vector<int64_t>& handles, int64_t width) const {
    std::vector<int32_t> out(handles.size() * static_cast<size_t>(width), 0);
    for (size_t i = 0; i < handles.size(); ++i) {
      const SchedRequest& r = reqs_.at(handles[i]);
      if (static_cast<int64_t>(r.blocks.size()) > width) throw std::out_of_range("block table narrower than the request's page list");
      for (size_t j = 0; j < r.blocks.size(); ++j) out[i * width + j] = static_cast<int32_t>(r.blocks[j]);
    }
    return out;
  }

  // Slot compaction (running sequences move to the lowest slots so a smaller decode-graph bucket serves them):
  // moves[i] = (handle, destination slot); the vacated slots return to the free list, which is kept sorted high -> low.
  void move_slots(const std::vector<std::pair<int64_t, int64_t>>& moves) {
    for (auto& mv : moves) {
      SchedRequest& r = reqs_.at(mv.first);
      auto it = std::find(free_slots_.begin(), free_slots_.end(), mv.second);
      if (it == free_slots_.end()) throw std::invalid_argument("move_slots: destination slot is not free");
      *it = r.slot;
      r.slot = mv.second;
    }
    std::sort(free_slots_.begin(), free_slots_.end(), [](int64_t a, int64_t b) { return a > b; });
  }
  std::vector<int64_t> free_slots() const { return free_slots_; }

  // Release a request's slot and pages.  `tokens` = prompt + generated tokens; with `to_cache` every FULL page whose KV is
  // final (all tokens except the last generated one, which was never fed back) enters the prefix cache first.
  void finish(int64_t h, const std::vector<int64_t>& tokens, bool to_cache) {
    auto it = reqs_.find(h);
    if (it == reqs_.end()) return;
    SchedRequest& r = it->second;
    if (r.live) {
      free_slots_.push_back(r.slot);
      if (!r.blocks.empty()) {
        if (to_cache && prefix_cache_ && r.cacheable) {
          const int64_t n = static_cast<int64_t>(tokens.size());
          const int64_t n_full = (n > 0 ? n - 1 : 0) / page_;
          std::vector<int64_t> head(r.blocks.begin(), r.blocks.begin() + std::min<int64_t>(n_full, static_cast<int64_t>(r.blocks.size())));
          kv_->insert(kv_->page_keys(tokens, n_full), head);
        }
        kv_->release(r.blocks);
      }
    }
    reqs_.erase(it);
  }

 private:
  KvPageManager* kv_;
  int64_t page_;
  bool prefix_cache_;
  int64_t next_ = 0;
  std::unordered_map<int64_t, SchedRequest> reqs_;
  std::deque<int64_t> waiting_;
  std::vector<int64_t> free_slots_;
};

}  // namespace prl
