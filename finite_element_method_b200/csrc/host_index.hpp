// Host-side indices for the reference's duplicate checks (same node coordinates, same element node
// set), built for bulk loading: 256 independent open-addressing shards so a batch of millions of
// keys can be inserted by all host cores at once (each thread owns a subset of the shards and scans
// the batch's precomputed hashes), while single adds stay O(1). The reference does a linear scan
// per add (methods_for_node_data_handle.rs:52-62, methods_for_truss_data_handle.rs:32-47, ...).
#pragma once
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

namespace femgpu {

inline unsigned host_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return std::max(1u, std::min(n ? n : 1u, 32u));
}

// run fn(t, n_threads) on n_threads threads (inline when n_threads == 1)
template <typename F>
inline void parallel_run(unsigned n_threads, F&& fn) {
  if (n_threads <= 1) {
    fn(0u, 1u);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(n_threads - 1);
  for (unsigned t = 1; t < n_threads; ++t) th.emplace_back([&, t] { fn(t, n_threads); });
  fn(0u, n_threads);
  for (auto& x : th) x.join();
}

template <typename F>
inline void parallel_chunks(size_t n, size_t min_per_thread, F&& fn /* (begin, end) */) {
  unsigned T = unsigned(std::min<size_t>(host_threads(), std::max<size_t>(1, n / std::max<size_t>(1, min_per_thread))));
  parallel_run(T, [&](unsigned t, unsigned nt) {
    size_t b = n * t / nt, e = n * (t + 1) / nt;
    if (b < e) fn(b, e);
  });
}

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// (hash, id) multiset keyed by a 64-bit hash; the caller supplies equality on ids for the rare case
// of two different keys with one hash. id 0xFFFFFFFF marks an empty slot, 0xFFFFFFFE a tombstone.
class ShardedIndex {
 public:
  static constexpr unsigned kShards = 256;
  static constexpr uint32_t kEmpty = 0xFFFFFFFFu, kDead = 0xFFFFFFFEu;

  // forget the entries, keep the tables (FEM::reset of a re-used instance: no fresh pages to fault in)
  void clear() {
    parallel_run(std::min(host_threads(), 8u), [&](unsigned t, unsigned nt) {
      for (unsigned sh = t; sh < kShards; sh += nt) {
        Shard& s = shards_[sh];
        if (s.used) std::fill(s.id.begin(), s.id.end(), kEmpty);
        s.used = 0;
      }
    });
  }

  // Looks for an entry equal to (hash, probe) under `same(existing_id)`; returns its id or kEmpty.
  template <typename Same>
  uint32_t find(uint64_t hash, Same&& same) const {
    const Shard& s = shards_[hash >> 56];
    if (s.cap == 0) return kEmpty;
    size_t mask = s.cap - 1, i = size_t(hash) & mask;
    for (;;) {
      uint32_t id = s.id[i];
      if (id == kEmpty) return kEmpty;
      if (id != kDead && s.hash[i] == hash && same(id)) return id;
      i = (i + 1) & mask;
    }
  }

  // insert without duplicate check (the caller has called find)
  void insert(uint64_t hash, uint32_t id) {
    Shard& s = shards_[hash >> 56];
    if ((s.used + 1) * 2 > s.cap) grow(s, std::max<size_t>(16, s.cap * 2));
    place(s, hash, id);
  }

  template <typename Same>
  void erase(uint64_t hash, uint32_t id_to_erase, Same&&) {
    Shard& s = shards_[hash >> 56];
    if (s.cap == 0) return;
    size_t mask = s.cap - 1, i = size_t(hash) & mask;
    for (;;) {
      uint32_t id = s.id[i];
      if (id == kEmpty) return;
      if (id == id_to_erase && s.hash[i] == hash) {
        s.id[i] = kDead;
        return;
      }
      i = (i + 1) & mask;
    }
  }

  // Bulk: hashes[i] belongs to new id first_id + i. For every i (ascending inside each shard) looks
  // for an equal earlier entry; if none, inserts. Returns the smallest i that found a duplicate
  // (n if none). `same(existing_id, i)` decides equality. Runs on all host cores.
  template <typename Same>
  size_t insert_batch(const uint64_t* hashes, size_t n, uint32_t first_id, Same&& same) {
    unsigned T = n >= 32768 ? std::min(host_threads(), kShards) : 1u;
    // reserve: expected entries per shard (uniform hash) with slack, so no shard grows mid-batch often
    size_t per = n / kShards + n / (kShards * 4) + 16;
    std::atomic<size_t> first_dup(n);
    parallel_run(T, [&](unsigned t, unsigned nt) {
      for (unsigned sh = t; sh < kShards; sh += nt) {
        Shard& s = shards_[sh];
        size_t want = (s.used + per) * 2;
        if (want > s.cap) {
          size_t cap = 16;
          while (cap < want) cap <<= 1;
          grow(s, cap);
        }
      }
      size_t local_first = n;
      for (size_t i = 0; i < n; ++i) {
        uint64_t hsh = hashes[i];
        unsigned sh = unsigned(hsh >> 56);
        if (sh % nt != t) continue;
        Shard& s = shards_[sh];
        size_t mask = s.cap - 1, p = size_t(hsh) & mask;
        bool dup = false;
        for (;;) {
          uint32_t id = s.id[p];
          if (id == kEmpty) break;
          if (id != kDead && s.hash[p] == hsh && same(id, i)) {
            dup = true;
            break;
          }
          p = (p + 1) & mask;
        }
        if (dup) {
          if (i < local_first) local_first = i;
          continue;  // a duplicate is never inserted
        }
        if ((s.used + 1) * 2 > s.cap) grow(s, s.cap * 2);
        place(s, hsh, first_id + uint32_t(i));
      }
      size_t cur = first_dup.load();
      while (local_first < cur && !first_dup.compare_exchange_weak(cur, local_first)) {
      }
    });
    return first_dup.load();
  }

  // remove the batch entries with index >= from (after a failed batch)
  void erase_batch(const uint64_t* hashes, size_t from, size_t n, uint32_t first_id) {
    for (size_t i = from; i < n; ++i) erase(hashes[i], first_id + uint32_t(i), [](uint32_t) { return true; });
  }

 private:
  struct Shard {
    std::vector<uint64_t> hash;
    std::vector<uint32_t> id;
    size_t cap = 0, used = 0;
  };
  static void place(Shard& s, uint64_t hash, uint32_t id) {
    size_t mask = s.cap - 1, i = size_t(hash) & mask;
    while (s.id[i] != kEmpty) i = (i + 1) & mask;  // tombstones are not reused; tables only grow
    s.hash[i] = hash;
    s.id[i] = id;
    ++s.used;
  }
  static void grow(Shard& s, size_t cap) {
    Shard n;
    n.cap = cap;
    n.hash.assign(cap, 0);
    n.id.assign(cap, kEmpty);
    for (size_t i = 0; i < s.cap; ++i)
      if (s.id[i] != kEmpty && s.id[i] != kDead) place(n, s.hash[i], s.id[i]);
    s = std::move(n);
  }
  Shard shards_[kShards];
};

}  // namespace femgpu
