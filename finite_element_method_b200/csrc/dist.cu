// Multi-GPU interface exchange: one process per GPU, one handle per process, joined by NCCL.
//
// Partition: rank g owns the contiguous node-index range [own_begin, own_end) and
// therefore the CSR rows [6*own_begin, 6*own_end); it is given every node but only the elements
// whose lowest-index node it owns. Its assembly kernel also produces partial sums for rows it does
// not own ("ghost rows", contiguous in its local CSR because rows are ordered by node index).
//   symbolic  the ghost node-pair block keys are sent to the owning ranks (ncclSend/ncclRecv), which
//             merge them into their pattern so every remote contribution has a slot
//   numeric   ghost blocks are packed (36 doubles each) and exchanged with one grouped
//             ncclSend/ncclRecv per neighbour on the handle's stream; the owner adds the received
//             partials one source rank at a time, in rank order -> deterministic
// NCCL is loaded with dlopen at femgpu_dist_init() so the library has no link-time dependency on it
// (and shares the copy a host framework such as PyTorch may already have loaded).
#include <dlfcn.h>

#include <cstring>

#include "common.cuh"

namespace femgpu {

namespace {

// minimal slice of nccl.h (ABI-stable since NCCL 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt64 = 4, ncclUint64 = 5, ncclFloat64 = 8 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* nccl() {
  static NcclApi api;
  if (api.lib || !api.error.empty()) return &api;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    api.error = std::string("could not dlopen libnccl.so.2: ") + dlerror();
    return &api;
  }
  auto sym = [&](const char* n) {
    void* p = dlsym(api.lib, n);
    if (!p && api.error.empty()) api.error = std::string("missing NCCL symbol ") + n;
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  return &api;
}

#define NCCL_CHECK(h, expr)                                                                   \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != 0)                                                                              \
      return (h)->fail(FEMGPU_ERR_NCCL, std::string("NCCL error: ") + nccl()->GetErrorString(_r) + \
                                            " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

// one thread per ghost block: gather its 6x6 (3x3 zero-padded) values into the dense send buffer
__global__ void pack_ghost_kernel(uint32_t n, uint32_t first_block, int key_bits,
                                  const uint64_t* __restrict__ blk_key,
                                  const uint32_t* __restrict__ blk_off,
                                  const uint32_t* __restrict__ node_len,
                                  const int64_t* __restrict__ node_base,
                                  const double* __restrict__ values, double* __restrict__ out) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint32_t blk = first_block + t;
  uint32_t a = uint32_t(blk_key[blk] >> key_bits);
  uint32_t l03 = node_len[2 * a], l35 = node_len[2 * a + 1];
  uint32_t o03 = blk_off[2 * blk], o35 = blk_off[2 * blk + 1];
  bool full = o35 != 0xFFFFFFFFu;
  int64_t base = node_base[a];
  double* o = out + size_t(t) * 36;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 6; ++j) {
      o[6 * i + j] = (full || j < 3) ? values[base + int64_t(i) * l03 + o03 + j] : 0.0;
      o[6 * (i + 3) + j] = full ? values[base + 3 * int64_t(l03) + int64_t(i) * l35 + o35 + j] : 0.0;
    }
}

// one thread per received block: add the partial block into the owner's CSR values
__global__ void apply_ghost_kernel(uint32_t n, const uint32_t* __restrict__ dst_block,
                                   const uint8_t* __restrict__ src_full, int key_bits,
                                   const uint64_t* __restrict__ blk_key,
                                   const uint32_t* __restrict__ blk_off,
                                   const uint32_t* __restrict__ node_len,
                                   const int64_t* __restrict__ node_base,
                                   const double* __restrict__ in, double* __restrict__ values) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  uint32_t blk = dst_block[t];
  uint32_t a = uint32_t(blk_key[blk] >> key_bits);
  uint32_t l03 = node_len[2 * a], l35 = node_len[2 * a + 1];
  uint32_t o03 = blk_off[2 * blk], o35 = blk_off[2 * blk + 1];
  bool full = src_full[t] != 0;  // the owner's block is full whenever any source's is
  int64_t base = node_base[a];
  const double* p = in + size_t(t) * 36;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < (full ? 6 : 3); ++j) values[base + int64_t(i) * l03 + o03 + j] += p[6 * i + j];
  if (full)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 6; ++j)
        values[base + 3 * int64_t(l03) + int64_t(i) * l35 + o35 + j] += p[6 * (i + 3) + j];
}

}  // namespace

// ---- helpers used by symbolic.cu ------------------------------------------------------------------

// every rank contributes `n` int64 (host) and receives world*n (host)
int32_t dist_allgather_i64(Handle* h, const int64_t* send, int64_t* recv, size_t n) {
  DistState& D = h->dist;
  Handle::DistScratch& S = h->dist_scratch;
  FEMGPU_CUDA_CHECK(h, S.i64.reserve((size_t(D.world) + 1) * n));
  int64_t* d_send = S.i64.p;
  int64_t* d_recv = S.i64.p + n;
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(d_send, send, n * 8, cudaMemcpyHostToDevice, h->stream));
  NCCL_CHECK(h, nccl()->AllGather(d_send, d_recv, n, ncclInt64, (ncclComm_t)D.comm, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaMemcpyAsync(recv, d_recv, size_t(D.world) * n * 8, cudaMemcpyDeviceToHost, h->stream));
  FEMGPU_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
  return 0;
}

// personalised exchange of 8-byte items: send_counts[r] items at send + send_offs[r] go to rank r
int32_t dist_exchange_8(Handle* h, const void* send, const int64_t* send_offs, const int64_t* send_counts,
                        void* recv, const int64_t* recv_offs, const int64_t* recv_counts) {
  DistState& D = h->dist;
  NCCL_CHECK(h, nccl()->GroupStart());
  for (int r = 0; r < D.world; ++r) {
    if (r == D.rank) continue;
    if (send_counts[r])
      NCCL_CHECK(h, nccl()->Send(static_cast<const char*>(send) + send_offs[r] * 8, size_t(send_counts[r]), ncclUint64,
                                 r, (ncclComm_t)D.comm, h->stream));
    if (recv_counts[r])
      NCCL_CHECK(h, nccl()->Recv(static_cast<char*>(recv) + recv_offs[r] * 8, size_t(recv_counts[r]), ncclUint64, r,
                                 (ncclComm_t)D.comm, h->stream));
  }
  NCCL_CHECK(h, nccl()->GroupEnd());
  return 0;
}

int32_t dist_numeric_exchange(Handle* h) {
  DistState& D = h->dist;
  if (!D.enabled) return 0;
  const int W = D.world;
  D.last_sent = D.last_recv = 0;
  // pack ghost blocks, one dense run per destination rank
  for (int r = 0; r < W; ++r) {
    if (r == D.rank || D.send_blocks[r] == 0) continue;
    uint32_t n = uint32_t(D.send_blocks[r]);
    pack_ghost_kernel<<<div_up(n, 128), 128, 0, h->stream>>>(n, uint32_t(D.send_first_block[r]), h->key_bits,
                                                             h->blk_key.p, h->blk_off.p, h->node_len.p,
                                                             h->node_base.p, h->values.p,
                                                             D.send_buf.p + size_t(D.send_off[r]) * 36);
    h->launches++;
  }
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  NCCL_CHECK(h, nccl()->GroupStart());
  for (int r = 0; r < W; ++r) {
    if (r == D.rank) continue;
    if (D.send_blocks[r]) {
      NCCL_CHECK(h, nccl()->Send(D.send_buf.p + size_t(D.send_off[r]) * 36, size_t(D.send_blocks[r]) * 36, ncclFloat64,
                                 r, (ncclComm_t)D.comm, h->stream));
      D.last_sent += uint64_t(D.send_blocks[r]) * 36 * 8;
    }
    if (D.recv_blocks[r]) {
      NCCL_CHECK(h, nccl()->Recv(D.recv_buf.p + size_t(D.recv_off[r]) * 36, size_t(D.recv_blocks[r]) * 36, ncclFloat64,
                                 r, (ncclComm_t)D.comm, h->stream));
      D.last_recv += uint64_t(D.recv_blocks[r]) * 36 * 8;
    }
  }
  NCCL_CHECK(h, nccl()->GroupEnd());
  // owner adds the partials, one source rank after the other in rank order (fixed order)
  for (int r = 0; r < W; ++r) {
    if (r == D.rank || D.recv_blocks[r] == 0) continue;
    uint32_t n = uint32_t(D.recv_blocks[r]);
    apply_ghost_kernel<<<div_up(n, 128), 128, 0, h->stream>>>(
        n, D.recv_dst_block.p + D.recv_off[r], D.recv_full.p + D.recv_off[r], h->key_bits, h->blk_key.p,
        h->blk_off.p, h->node_len.p, h->node_base.p, D.recv_buf.p + size_t(D.recv_off[r]) * 36, h->values.p);
    h->launches++;
  }
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  return 0;
}

void dist_destroy(Handle* h) {
  if (h->dist.comm && nccl()->CommDestroy) nccl()->CommDestroy((ncclComm_t)h->dist.comm);
  h->dist.comm = nullptr;
  h->dist.enabled = false;
  h->dist.send_buf.release();
  h->dist.recv_buf.release();
  h->dist.recv_dst_block.release();
  h->dist.recv_full.release();
  h->dist.remote_keys.release();
  h->dist_scratch.i64.release();
}

}  // namespace femgpu

using namespace femgpu;

extern "C" {

int32_t femgpu_dist_unique_id(uint8_t out[128]) {
  if (!out) return FEMGPU_ERR_USAGE;
  NcclApi* api = nccl();
  if (!api->error.empty()) return FEMGPU_ERR_NCCL;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != 0) return FEMGPU_ERR_NCCL;
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  std::memcpy(out, &id, 128);
  return 0;
}

int32_t femgpu_dist_init(femgpu_t* h, int32_t rank, int32_t world, const uint8_t nccl_unique_id[128]) {
  if (!h || !nccl_unique_id) return FEMGPU_ERR_USAGE;
  if (h->device < 0) return h->fail(FEMGPU_ERR_NO_DEVICE, "staging-only handle");
  if (world < 1 || rank < 0 || rank >= world) return h->fail(FEMGPU_ERR_USAGE, "bad rank / world");
  NcclApi* api = nccl();
  if (!api->error.empty()) return h->fail(FEMGPU_ERR_NCCL, api->error);
  FEMGPU_CUDA_CHECK(h, cudaSetDevice(h->device));
  ncclUniqueId id;
  std::memcpy(&id, nccl_unique_id, 128);
  ncclComm_t comm = nullptr;
  NCCL_CHECK(h, api->CommInitRank(&comm, world, id, rank));
  DistState& D = h->dist;
  D.comm = comm;
  D.rank = rank;
  D.world = world;
  D.enabled = world > 1;
  D.send_blocks.assign(world, 0);
  D.recv_blocks.assign(world, 0);
  D.send_off.assign(world + 1, 0);
  D.recv_off.assign(world + 1, 0);
  D.send_first_block.assign(world, 0);
  h->symbolic_valid = false;
  h->values_valid = false;
  return 0;
}

int32_t femgpu_dist_set_ownership(femgpu_t* h, uint32_t node_index_begin, uint32_t node_index_end) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (node_index_begin > node_index_end || node_index_end > h->nodes_number)
    return h->fail(FEMGPU_ERR_USAGE, "ownership range outside [0, nodes_number]");
  h->dist.own_begin = node_index_begin;
  h->dist.own_end = node_index_end;
  h->dist.ownership_set = true;
  h->symbolic_valid = false;
  h->values_valid = false;
  return 0;
}

int32_t femgpu_dist_last_exchange_bytes(femgpu_t* h, uint64_t* sent, uint64_t* received) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (sent) *sent = h->dist.last_sent;
  if (received) *received = h->dist.last_recv;
  return 0;
}

}  // extern "C"
