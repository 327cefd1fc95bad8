// Host-side indices for the reference's duplicate checks (same node coordinates, same element node
// set), built for bulk loading: 1024 independent open-addressing shards so a batch of millions of
// keys can be inserted by all host cores at once (each thread owns a subset of the shards and scans
// the batch's precomputed hashes), while single adds stay O(1). The reference does a linear scan
// per add (methods_for_node_data_handle.rs:52-62, methods_for_truss_data_handle.rs:32-47, ...).
#pragma once
#include <stdint.h>
#include <pthread.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <type_traits>
#include <functional>
#include <thread>
#include <vector>

namespace femgpu {

// host threads of the bulk paths: the cores of the machine, at most 32 (FEMGPU_HOST_THREADS overrides), divided by
// the number of ranks that share the machine once a handle joins a multi-rank run (host_threads_share: every rank
// runs its own batched adds at the same time; 8 ranks x 16 threads on 16 cores only wait for each other)
inline std::atomic<unsigned>& host_threads_divisor() {
  static std::atomic<unsigned> d(1);
  return d;
}
inline unsigned host_threads() {
  static const int forced = [] {
    if (const char* e = getenv("FEMGPU_HOST_THREADS")) {
      int v = atoi(e);
      if (v > 0) return std::min(v, 256);
    }
    return 0;
  }();
  if (forced) return unsigned(forced);
  static const unsigned cores = [] {
    unsigned n = std::thread::hardware_concurrency();
    return std::max(1u, std::min(n ? n : 1u, 32u));
  }();
  const unsigned d = std::max(1u, host_threads_divisor().load(std::memory_order_relaxed));
  return std::max(1u, (cores + d - 1) / d);
}
inline void host_threads_share(unsigned ranks_on_this_machine) {
  host_threads_divisor().store(std::max(1u, ranks_on_this_machine), std::memory_order_relaxed);
}

// Worker threads of the bulk paths, started once per process and parked on a condition variable between
// regions: a batched add runs a dozen short parallel regions, and starting 31 threads for each of them costs
// more than some of the regions themselves.
class HostPool {
 public:
  // never destroyed: at process exit the parked workers simply go with the process (no join against threads that a
  // fork()ed child does not have); after a fork the child starts threads per region instead of using the pool
  static HostPool& get() {
    static HostPool* pool = [] {
      pthread_atfork(nullptr, nullptr, [] { forked() = true; });
      return new HostPool;
    }();
    return *pool;
  }
  // fn(t, n_threads) for t = 0 .. n_threads-1, t = 0 on the calling thread; returns when all are done.
  // A region entered while another one is running (a second handle on another user thread, or a nested call)
  // starts its own threads instead of waiting for the pool.
  template <typename F>
  void run(unsigned n_threads, F&& fn) {
    std::unique_lock<std::mutex> region(region_mutex_, std::defer_lock);
    if (!forked()) (void)region.try_lock();
    if (!region.owns_lock() || inside()) {
      std::vector<std::thread> th;
      th.reserve(n_threads - 1);
      for (unsigned t = 1; t < n_threads; ++t) th.emplace_back([&, t] { fn(t, n_threads); });
      fn(0u, n_threads);
      for (auto& x : th) x.join();
      return;
    }
    grow(n_threads - 1);
    {
      std::lock_guard<std::mutex> lk(m_);
      call_ = [](void* ctx, unsigned t, unsigned nt) { (*static_cast<typename std::remove_reference<F>::type*>(ctx))(t, nt); };
      ctx_ = const_cast<void*>(static_cast<const void*>(&fn));
      n_threads_ = n_threads;
      remaining_ = n_threads - 1;
      ++generation_;
    }
    wake_.notify_all();
    inside() = true;
    fn(0u, n_threads);
    inside() = false;
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [&] { return remaining_ == 0; });
  }

 private:
  HostPool() = default;
  static bool& forked() {
    static bool f = false;
    return f;
  }
  static bool& inside() {
    static thread_local bool in = false;
    return in;
  }
  void grow(unsigned want) {
    while (workers_.size() < want) {
      const unsigned id = unsigned(workers_.size()) + 1;  // thread index inside a region
      uint64_t seen;
      {
        std::lock_guard<std::mutex> lk(m_);
        seen = generation_;
      }
      workers_.emplace_back([this, id, seen]() mutable {
        inside() = true;
        for (;;) {
          void (*call)(void*, unsigned, unsigned);
          void* ctx;
          unsigned nt;
          {
            std::unique_lock<std::mutex> lk(m_);
            wake_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            call = call_; ctx = ctx_; nt = n_threads_;
          }
          if (id < nt) {
            call(ctx, id, nt);
            std::lock_guard<std::mutex> lk(m_);
            if (--remaining_ == 0) done_.notify_one();
          }
        }
      });
    }
  }
  std::mutex region_mutex_, m_;
  std::condition_variable wake_, done_;
  std::vector<std::thread> workers_;
  void (*call_)(void*, unsigned, unsigned) = nullptr;
  void* ctx_ = nullptr;
  unsigned n_threads_ = 0, remaining_ = 0;
  uint64_t generation_ = 0;
  bool stop_ = false;
};

// run fn(t, n_threads) on n_threads threads (inline when n_threads == 1)
template <typename F>
inline void parallel_run(unsigned n_threads, F&& fn) {
  if (n_threads <= 1) {
    fn(0u, 1u);
    return;
  }
  HostPool::get().run(n_threads, fn);
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
// of two different keys with one hash. A slot is one 64-bit word: bits 22..53 of the hash in the upper half
// (they give the slot inside the shard and the quick comparison — the top ten bits chose the shard), the id in the
// lower half; id 0xFFFFFFFF marks an empty slot, 0xFFFFFFFE a tombstone. One cache line per probe.
class ShardedIndex {
 public:
  static constexpr unsigned kShift = 54, kShards = 1u << (64 - kShift);  // 1024 shards, keyed by the top hash bits
  static constexpr uint32_t kEmpty = 0xFFFFFFFFu, kDead = 0xFFFFFFFEu;

  // forget the entries, keep the tables (FEM::reset of a re-used instance: no fresh pages to fault in)
  void clear() {
    parallel_run(host_threads(), [&](unsigned t, unsigned nt) {
      for (unsigned sh = t; sh < kShards; sh += nt) {
        Shard& s = shards_[sh];
        if (s.used) std::fill(s.slot.begin(), s.slot.end(), ~0ull);
        s.used = 0;
      }
    });
  }

  // Looks for an entry equal to (hash, probe) under `same(existing_id)`; returns its id or kEmpty.
  template <typename Same>
  uint32_t find(uint64_t hash, Same&& same) const {
    const Shard& s = shards_[hash >> kShift];
    if (s.cap == 0) return kEmpty;
    const uint32_t tag = tag_of(hash);
    size_t mask = s.cap - 1, i = tag & mask;
    for (;;) {
      const uint64_t w = s.slot[i];
      const uint32_t id = uint32_t(w);
      if (id == kEmpty) return kEmpty;
      if (id != kDead && uint32_t(w >> 32) == tag && same(id)) return id;
      i = (i + 1) & mask;
    }
  }

  // insert without duplicate check (the caller has called find)
  void insert(uint64_t hash, uint32_t id) {
    Shard& s = shards_[hash >> kShift];
    if ((s.used + 1) * 2 > s.cap) grow(s, std::max<size_t>(16, s.cap * 2));
    place(s, tag_of(hash), id);
  }

  template <typename Same>
  void erase(uint64_t hash, uint32_t id_to_erase, Same&&) {
    Shard& s = shards_[hash >> kShift];
    if (s.cap == 0) return;
    size_t mask = s.cap - 1, i = tag_of(hash) & mask;
    for (;;) {
      const uint32_t id = uint32_t(s.slot[i]);
      if (id == kEmpty) return;
      if (id == id_to_erase) {  // ids are unique
        s.slot[i] = (s.slot[i] & 0xFFFFFFFF00000000ull) | kDead;
        return;
      }
      i = (i + 1) & mask;
    }
  }

  // Bulk: hashes[i] belongs to new id first_id + i. For every i (ascending inside each shard) looks
  // for an equal earlier entry; if none, inserts. Returns the smallest i that found a duplicate
  // (n if none). `same(existing_id, i)` decides equality. Runs on all host cores: the batch is first
  // partitioned by shard (histogram + scatter of (hash, i) pairs, order kept), then every shard is
  // filled by one thread while its table — a hundred KB or so — stays in that core's cache. Probing the
  // tables in batch order instead costs a cache-line miss per key and is bound by memory traffic.
  struct Item {
    uint32_t tag;
    uint32_t i;
  };
  template <typename Same>
  size_t insert_batch(const uint64_t* hashes, size_t n, uint32_t first_id, Same&& same, std::vector<Item>& scratch) {
    if (n < 4096) {  // single adds and small batches: probe in batch order
      for (size_t i = 0; i < n; ++i) {
        if (find(hashes[i], [&](uint32_t id) { return same(id, i); }) != kEmpty) return i;
        insert(hashes[i], first_id + uint32_t(i));
      }
      return n;
    }
    const unsigned T = n >= 32768 ? host_threads() : 1u;
    std::vector<uint32_t> hist(size_t(T) * kShards, 0u);
    parallel_run(T, [&](unsigned t, unsigned nt) {
      uint32_t* hg = hist.data() + size_t(t) * kShards;
      for (size_t i = n * t / nt, e = n * (t + 1) / nt; i < e; ++i) hg[hashes[i] >> kShift]++;
    });
    std::vector<uint32_t> shard_begin(kShards + 1);
    uint32_t running = 0;
    for (unsigned sh = 0; sh < kShards; ++sh) {
      shard_begin[sh] = running;
      for (unsigned t = 0; t < T; ++t) {
        uint32_t c = hist[size_t(t) * kShards + sh];
        hist[size_t(t) * kShards + sh] = running;
        running += c;
      }
    }
    shard_begin[kShards] = running;
    if (scratch.size() < n) scratch.resize(n);
    Item* items = scratch.data();
    parallel_run(T, [&](unsigned t, unsigned nt) {
      uint32_t* at = hist.data() + size_t(t) * kShards;
      for (size_t i = n * t / nt, e = n * (t + 1) / nt; i < e; ++i)
        items[at[hashes[i] >> kShift]++] = Item{tag_of(hashes[i]), uint32_t(i)};
    });
    std::atomic<size_t> first_dup(n);
    std::atomic<unsigned> next(0);
    parallel_run(T, [&](unsigned, unsigned) {
      size_t local_first = n;
      for (unsigned sh = next.fetch_add(1); sh < kShards; sh = next.fetch_add(1)) {
        const size_t b = shard_begin[sh], e = shard_begin[sh + 1];
        if (b == e) continue;
        Shard& s = shards_[sh];
        const size_t want = (s.used + (e - b)) * 2;
        if (want > s.cap) {
          size_t cap = 16;
          while (cap < want) cap <<= 1;
          grow(s, cap);
        }
        const size_t mask = s.cap - 1;
        uint64_t* slot = s.slot.data();
        constexpr size_t kAhead = 8;
        for (size_t k = b; k < e; ++k) {
          if (k + kAhead < e) __builtin_prefetch(slot + (items[k + kAhead].tag & mask), 1);
          const uint32_t tag = items[k].tag;
          const size_t i = items[k].i;
          size_t p = tag & mask;
          bool dup = false;
          for (;;) {
            const uint64_t w = slot[p];
            const uint32_t id = uint32_t(w);
            if (id == kEmpty) break;
            if (id != kDead && uint32_t(w >> 32) == tag && same(id, i)) {
              dup = true;
              break;
            }
            p = (p + 1) & mask;
          }
          if (dup) {
            if (i < local_first) local_first = i;
            continue;  // a duplicate is never inserted
          }
          slot[p] = (uint64_t(tag) << 32) | (first_id + uint32_t(i));  // the probe ended on the first empty slot of the chain
          ++s.used;
        }
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
    std::vector<uint64_t> slot;  // tag << 32 | id
    size_t cap = 0, used = 0;
  };
  static uint32_t tag_of(uint64_t hash) { return uint32_t(hash >> 22); }
  static void place(Shard& s, uint32_t tag, uint32_t id) {
    size_t mask = s.cap - 1, i = tag & mask;
    while (uint32_t(s.slot[i]) != kEmpty) i = (i + 1) & mask;  // tombstones are not reused; tables only grow
    s.slot[i] = (uint64_t(tag) << 32) | id;
    ++s.used;
  }
  static void grow(Shard& s, size_t cap) {
    Shard n;
    n.cap = cap;
    n.slot.assign(cap, ~0ull);
    for (size_t i = 0; i < s.cap; ++i) {
      const uint32_t id = uint32_t(s.slot[i]);
      if (id != kEmpty && id != kDead) place(n, uint32_t(s.slot[i] >> 32), id);
    }
    s = std::move(n);
  }
  Shard shards_[kShards];
};

}  // namespace femgpu
