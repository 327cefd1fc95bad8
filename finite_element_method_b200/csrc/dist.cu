// Multi-GPU interface exchange (one process per GPU). Placeholder until the NCCL path lands.
#include "common.cuh"

namespace femgpu {
int32_t dist_symbolic_exchange(Handle* h) { return h->fail(FEMGPU_ERR_USAGE, "dist not built"); }
int32_t dist_numeric_exchange(Handle* h) { return h->fail(FEMGPU_ERR_USAGE, "dist not built"); }
void dist_destroy(Handle*) {}
}  // namespace femgpu

extern "C" {
int32_t femgpu_dist_unique_id(uint8_t*) { return FEMGPU_ERR_USAGE; }
int32_t femgpu_dist_init(femgpu_t* h, int32_t, int32_t, const uint8_t*) { return h ? h->fail(FEMGPU_ERR_USAGE, "dist not built") : FEMGPU_ERR_USAGE; }
int32_t femgpu_dist_set_ownership(femgpu_t* h, uint32_t, uint32_t) { return h ? h->fail(FEMGPU_ERR_USAGE, "dist not built") : FEMGPU_ERR_USAGE; }
int32_t femgpu_dist_last_exchange_bytes(femgpu_t*, uint64_t*, uint64_t*) { return FEMGPU_ERR_USAGE; }
}
