// ThreadSanitizer harness of the parallel host structures of the batched adds (csrc/common.cuh NumberMap::insert_batch,
// csrc/host_index.hpp ShardedIndex::insert_batch, HostPool): labels in no order with a repeat placed in another
// thread's chunk, ascending labels on top, hashed keys with a repeat. Built and run by tests/test_host_threads.py.
#include "common.cuh"
#include <cstdio>
#include <random>
using namespace femgpu;
int main() {
  const size_t n = 300000;
  std::mt19937_64 rng(7);
  for (int trial = 0; trial < 4; ++trial) {
    std::vector<uint32_t> lab(n);
    for (size_t i = 0; i < n; ++i) lab[i] = uint32_t(i + 1);
    std::shuffle(lab.begin(), lab.end(), rng);
    size_t d1 = 100000 + rng() % 100000, d0 = rng() % d1;
    if (trial > 0) lab[d1] = lab[d0];
    NumberMap m;
    size_t stop = m.insert_batch(lab.data(), n, 0);
    size_t expect = trial > 0 ? d1 : n;
    uint32_t idx = 0;
    bool ok = stop == expect && m.count == expect && m.find(lab[0], &idx) && idx == 0;
    if (trial > 0) ok = ok && m.find(lab[d0], &idx) && idx == d0 && !m.find(lab[std::min(n - 1, d1 + 1)], &idx);
    printf("numbers trial %d: stop %zu expect %zu %s\n", trial, stop, expect, ok ? "ok" : "WRONG");
    // ascending batch on top
    std::vector<uint32_t> up(n);
    for (size_t i = 0; i < n; ++i) up[i] = uint32_t(n + 10 + i);
    size_t s2 = m.insert_batch(up.data(), n, uint32_t(n));
    printf("  ascending: %zu %s\n", s2, s2 == n ? "ok" : "WRONG");
    m.clear();
    std::vector<uint64_t> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = mix64(rng());
    if (trial > 0) h[d1] = h[d0];
    ShardedIndex ix;
    std::vector<ShardedIndex::Item> scratch;
    size_t dup = ix.insert_batch(h.data(), n, 0, [&](uint32_t a, size_t b) { return h[a] == h[b]; }, scratch);
    printf("index trial %d: dup %zu expect %zu %s\n", trial, dup, expect, dup == expect ? "ok" : "WRONG");
    ix.clear();
  }
  printf("HOST_STRUCTURES_DONE\n");
}
