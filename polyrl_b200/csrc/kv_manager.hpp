// polyrl_b200 — native KV page manager of the rollout engine: ref-counted page allocator + cross-request
// prefix cache (hash chain per full page, LRU eviction of pages only the cache still references).
//
// Host-only C++ (no CUDA): this is the scheduler-side bookkeeping that runs once per admitted / finished request
// on the decode loop's critical host path; with long prompts the Python version spends its time building
// per-page tuples and hashing them.  Same semantics as the Python reference implementation kept in
// rollout/engine.py (`BlockAllocator`, `PrefixCache`), which the CPU tests compare against op by op.
// Role in the reference stack: SGLang's token-to-KV pool allocator + radix cache
// (3rdparty/verl/verl/workers/rollout/sglang_rollout/sglang_rollout.py drives them through the engine).
#pragma once
#include <cstdint>
#include <deque>
#include <list>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace prl {

class KvPageManager {
 public:
  explicit KvPageManager(int64_t num_blocks, int64_t page_size) : num_blocks_(num_blocks), page_(page_size), ref_(num_blocks, 0) {
    if (num_blocks < 1 || page_size < 1) throw std::invalid_argument("KvPageManager: bad sizes");
    for (int64_t b = 1; b < num_blocks; ++b) free_.push_back(b);          // block 0 = scratch page of idle batch slots
  }

  // ---- allocator
  int64_t num_blocks() const { return num_blocks_; }
  int64_t available() const { return static_cast<int64_t>(free_.size()); }
  int64_t ref(int64_t b) const { return ref_.at(b); }

  std::vector<int64_t> alloc(int64_t n) {
    if (n > available()) throw std::runtime_error("KV cache exhausted");
    std::vector<int64_t> out;
    out.reserve(n);
    for (int64_t i = 0; i < n; ++i) {
      const int64_t b = free_.front();
      free_.pop_front();
      ref_[b] = 1;
      out.push_back(b);
    }
    return out;
  }
  void share(const std::vector<int64_t>& blocks) {
    for (int64_t b : blocks) ++ref_.at(b);
  }
  void release(const std::vector<int64_t>& blocks) {
    for (int64_t b : blocks) {
      if (--ref_.at(b) == 0) free_.push_back(b);
    }
  }

  // ---- prefix cache
  // key of page i = hash chain over every token of pages 0..i  (a hit on page i implies hits on 0..i-1)
  std::vector<uint64_t> page_keys(const std::vector<int64_t>& tokens, int64_t max_pages) const {
    std::vector<uint64_t> keys;
    if (max_pages <= 0) return keys;
    keys.reserve(max_pages);
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int64_t p = 0; p < max_pages; ++p) {
      for (int64_t i = p * page_; i < (p + 1) * page_; ++i) {
        const uint64_t t = i < static_cast<int64_t>(tokens.size()) ? static_cast<uint64_t>(tokens[i]) : ~0ull;
        h = mix(h ^ (t + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2)));
      }
      keys.push_back(h);
    }
    return keys;
  }

  std::vector<int64_t> match(const std::vector<uint64_t>& keys) {
    std::vector<int64_t> out;
    for (uint64_t k : keys) {
      auto it = by_key_.find(k);
      if (it == by_key_.end()) break;
      out.push_back(it->second);
    }
    for (int64_t b : out) touch(b);
    hits_ += static_cast<int64_t>(out.size());
    misses_ += static_cast<int64_t>(keys.size() - out.size());
    return out;
  }

  void insert(const std::vector<uint64_t>& keys, const std::vector<int64_t>& blocks) {
    const size_t n = keys.size() < blocks.size() ? keys.size() : blocks.size();
    for (size_t i = 0; i < n; ++i) {
      const uint64_t k = keys[i];
      const int64_t b = blocks[i];
      if (by_key_.count(k) || key_of_.count(b)) continue;
      by_key_[k] = b;
      key_of_[b] = k;
      ++ref_.at(b);                                   // the cache holds one reference
      lru_.push_back(b);
      lru_pos_[b] = std::prev(lru_.end());
    }
  }

  // free up to `want` pages that only the cache still references, oldest first; returns how many were freed
  int64_t evict(int64_t want) {
    int64_t freed = 0;
    for (auto it = lru_.begin(); it != lru_.end() && freed < want;) {
      const int64_t b = *it;
      ++it;                                            // drop() erases the node
      if (ref_[b] == 1) {
        drop(b);
        ++freed;
      }
    }
    return freed;
  }

  void flush() {
    std::vector<int64_t> all;
    all.reserve(key_of_.size());
    for (auto& kv : key_of_) all.push_back(kv.first);
    for (int64_t b : all) drop(b);
  }

  int64_t cached_pages() const { return static_cast<int64_t>(key_of_.size()); }
  int64_t hits() const { return hits_; }
  int64_t misses() const { return misses_; }

 private:
  static uint64_t mix(uint64_t x) {                    // splitmix64 finaliser
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
  }
  void touch(int64_t b) {
    auto it = lru_pos_.find(b);
    if (it == lru_pos_.end()) return;
    lru_.erase(it->second);
    lru_.push_back(b);
    it->second = std::prev(lru_.end());
  }
  void drop(int64_t b) {
    auto kit = key_of_.find(b);
    if (kit == key_of_.end()) return;
    by_key_.erase(kit->second);
    key_of_.erase(kit);
    auto it = lru_pos_.find(b);
    if (it != lru_pos_.end()) {
      lru_.erase(it->second);
      lru_pos_.erase(it);
    }
    if (--ref_.at(b) == 0) free_.push_back(b);
  }

  int64_t num_blocks_, page_;
  std::deque<int64_t> free_;
  std::vector<int32_t> ref_;
  std::unordered_map<uint64_t, int64_t> by_key_;
  std::unordered_map<int64_t, uint64_t> key_of_;
  std::list<int64_t> lru_;                             // oldest first
  std::unordered_map<int64_t, std::list<int64_t>::iterator> lru_pos_;
  int64_t hits_ = 0, misses_ = 0;
};

}  // namespace prl
