// Multi-GPU interface exchange: one process per GPU, one handle per process, joined by NCCL.
//
// Partition: rank g owns the contiguous node-index range [own_begin, own_end) and
// therefore the CSR rows [6*own_begin, 6*own_end); it is given every node but only the elements
// whose lowest-index node it owns. Its assembly kernel also produces partial sums for rows it does
// not own ("ghost rows", contiguous in its local CSR because rows are ordered by node index).
//   symbolic  the ghost node-pair block keys are sent to the owning ranks (ncclSend/ncclRecv), which
//             merge them into their pattern so every remote contribution has a slot
//   numeric   ghost blocks are packed (36 doubles each) by a kernel that stores them STRAIGHT INTO THE
//             OWNER'S HBM over NVLink (peer memory mapped with CUDA IPC) and then raises a flag there;
//             the owner's apply kernel waits for the flag and adds the partials one source rank at a
//             time, in rank order -> deterministic. No NCCL call, no host synchronisation and no
//             matched send/recv pair inside a numeric pass: a rank that runs a pass its neighbour never
//             runs gets FEMGPU_ERR_NCCL ("timed out") from the next synchronising call instead of a hung GPU.
//             When the peers cannot be mapped (no P2P, one process driving several handles) the
//             exchange falls back to one grouped ncclSend/ncclRecv per neighbour (FEMGPU_DIST_P2P=0 forces it).
// NCCL (symbolic-pass collectives, fallback exchange) is loaded with dlopen at femgpu_dist_init() so the
// library has no link-time dependency on it (and shares the copy a host framework such as PyTorch may
// already have loaded).
#include <dlfcn.h>
#include <unistd.h>

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


// ---- peer-to-peer exchange --------------------------------------------------------------------------
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// thread 0 of the CTA waits until *flag >= want (a flag in THIS device's memory, raised by a peer); gives up
// after timeout_ns and records `code` in the mapped host word so that the next synchronising call can report it
__device__ __forceinline__ void wait_flag(const uint64_t* flag, uint64_t want, uint64_t timeout_ns, uint32_t code,
                                          uint32_t* err) {
  if (threadIdx.x == 0 && ld_acquire_sys(flag) < want) {
    const uint64_t t0 = global_timer_ns();
    while (ld_acquire_sys(flag) < want) {
      __nanosleep(200);
      if (global_timer_ns() - t0 > timeout_ns) {
        atomicCAS(err, 0u, code);
        __threadfence_system();
        break;
      }
    }
  }
  __syncthreads();
}
// the last CTA of the launch to get here raises the peer's flag (after every CTA's peer accesses are fenced)
__device__ __forceinline__ void signal_when_all_done(uint32_t* counter, uint64_t* peer_flag, uint64_t value) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(counter, 1u) == gridDim.x - 1u) {
      *counter = 0u;  // ready for the next launch (stream-ordered after this one)
      __threadfence_system();
      st_release_sys(peer_flag, value);
    }
  }
}

// pack_ghost_kernel, writing into the owner's window: waits until the owner has consumed the pass that last
// used this ring slot, stores the blocks over NVLink, raises arrived[me] = epoch in the owner's window
__global__ void __launch_bounds__(128)
pack_ghost_p2p_kernel(uint32_t n, uint32_t first_block, int key_bits, const uint64_t* __restrict__ blk_key,
                      const uint32_t* __restrict__ blk_off, const uint32_t* __restrict__ node_len,
                      const int64_t* __restrict__ node_base, const double* __restrict__ values,
                      double* __restrict__ peer_out, const uint64_t* consumed_flag, uint64_t need_consumed,
                      uint64_t* peer_arrived_flag, uint64_t epoch, uint32_t* counter, uint64_t timeout_ns,
                      uint32_t err_code, uint32_t* err) {
  if (need_consumed) wait_flag(consumed_flag, need_consumed, timeout_ns, err_code, err);
  // one thread per (block, dof row): six values in, three 16-byte stores over NVLink out
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, b = t / 6u, row = t - 6u * b;
  if (b < n) {
    const uint32_t blk = first_block + b;
    const uint32_t a = uint32_t(blk_key[blk] >> key_bits);
    const uint32_t l03 = node_len[2 * a], l35 = node_len[2 * a + 1];
    const uint32_t o03 = blk_off[2 * blk], o35 = blk_off[2 * blk + 1];
    const bool full = o35 != 0xFFFFFFFFu;
    const int64_t base = node_base[a];
    double v[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (row < 3u) {
      const double* r = values + base + int64_t(row) * l03 + o03;
      for (int j = 0; j < (full ? 6 : 3); ++j) v[j] = r[j];
    } else if (full) {
      const double* r = values + base + 3 * int64_t(l03) + int64_t(row - 3u) * l35 + o35;
      for (int j = 0; j < 6; ++j) v[j] = r[j];
    }
    double2* o = reinterpret_cast<double2*>(peer_out + size_t(b) * 36 + 6u * row);
    o[0] = make_double2(v[0], v[1]);
    o[1] = make_double2(v[2], v[3]);
    o[2] = make_double2(v[4], v[5]);
  }
  signal_when_all_done(counter, peer_arrived_flag, epoch);
}

// apply_ghost_kernel on the ring slot the sender filled: waits for arrived[src] >= epoch (my memory), adds the
// partial blocks, raises consumed[me] = epoch in the sender's window
__global__ void __launch_bounds__(128)
apply_ghost_p2p_kernel(uint32_t n, const uint32_t* __restrict__ dst_block, const uint8_t* __restrict__ src_full,
                       int key_bits, const uint64_t* __restrict__ blk_key, const uint32_t* __restrict__ blk_off,
                       const uint32_t* __restrict__ node_len, const int64_t* __restrict__ node_base,
                       const double* in, double* __restrict__ values, const uint64_t* arrived_flag, uint64_t epoch,
                       uint64_t* peer_consumed_flag, uint32_t* counter, uint64_t timeout_ns, uint32_t err_code,
                       uint32_t* err) {
  wait_flag(arrived_flag, epoch, timeout_ns, err_code, err);
  // one thread per (received block, dof row)
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, b = t / 6u, row = t - 6u * b;
  if (b < n) {
    const uint32_t blk = dst_block[b];
    const uint32_t a = uint32_t(blk_key[blk] >> key_bits);
    const uint32_t l03 = node_len[2 * a], l35 = node_len[2 * a + 1];
    const uint32_t o03 = blk_off[2 * blk], o35 = blk_off[2 * blk + 1];
    const bool full = src_full[b] != 0;  // the owner's block is full whenever any source's is
    const int64_t base = node_base[a];
    // written by the peer: read through L2 (never a stale L1 line)
    const double2* p = reinterpret_cast<const double2*>(in + size_t(b) * 36 + 6u * row);
    if (row < 3u || full) {
      const double2 v0 = __ldcg(p), v1 = __ldcg(p + 1), v2 = __ldcg(p + 2);
      const double v[6] = {v0.x, v0.y, v1.x, v1.y, v2.x, v2.y};
      double* r = row < 3u ? values + base + int64_t(row) * l03 + o03
                           : values + base + 3 * int64_t(l03) + int64_t(row - 3u) * l35 + o35;
      for (int j = 0; j < (full ? 6 : 3); ++j) r[j] += v[j];
    }
  }
  signal_when_all_done(counter, peer_consumed_flag, epoch);
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

constexpr size_t kFlagStride = 64;  // bytes between two flags of a window
static inline uint64_t* win_arrived(unsigned char* win, int src) { return reinterpret_cast<uint64_t*>(win + kFlagStride * size_t(src)); }
static inline uint64_t* win_consumed(unsigned char* win, int world, int dst) {
  return reinterpret_cast<uint64_t*>(win + kFlagStride * size_t(world + dst));
}
static inline size_t win_flag_bytes(int world) { return (2 * kFlagStride * size_t(world) + 255) & ~size_t(255); }

static uint64_t p2p_timeout_ns() {
  static const uint64_t ns = [] {
    const char* q = getenv("FEMGPU_P2P_TIMEOUT_MS");
    const double ms = q ? atof(q) : 20000.0;
    return uint64_t((ms > 1.0 ? ms : 1.0) * 1e6);
  }();
  return ns;
}

// my ghost blocks -> the owners' windows (pass D.epoch), on stream `st`
static int32_t p2p_pack(Handle* h, cudaStream_t st) {
  DistState& D = h->dist;
  const int W = D.world;
  const uint64_t epoch = D.epoch;
  const uint32_t slot = uint32_t((epoch - 1) % DistState::kRing);
  const size_t flag_bytes = win_flag_bytes(W);
  for (int r = 0; r < W; ++r) {
    if (r == D.rank || D.send_blocks[r] == 0) continue;
    const uint32_t n = uint32_t(D.send_blocks[r]);
    // the ring slots of a window are sized by ITS OWNER's receive area
    double* out = reinterpret_cast<double*>(D.peer_win[r] + flag_bytes + size_t(slot) * D.peer_slot_bytes[r]) +
                  size_t(D.peer_recv_off[r]) * 36;
    const uint64_t need = epoch > uint64_t(DistState::kRing) ? epoch - DistState::kRing : 0;
    pack_ghost_p2p_kernel<<<div_up(uint64_t(n) * 6, 128), 128, 0, st>>>(
        n, uint32_t(D.send_first_block[r]), h->key_bits, h->blk_key.p, h->blk_off.p, h->node_len.p, h->node_base.p,
        h->values.p, out, win_consumed(D.win, W, r), need, win_arrived(D.peer_win[r], D.rank), epoch,
        D.done_count + r, p2p_timeout_ns(), 0x100u + uint32_t(r), D.d_err);
    h->launches++;
    D.last_sent += uint64_t(n) * 36 * 8;
  }
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  return 0;
}

// partials of the other ranks, in rank order (fixed order of the sums), on the handle's stream
static int32_t p2p_apply(Handle* h) {
  DistState& D = h->dist;
  const int W = D.world;
  const uint64_t epoch = D.epoch;
  const uint32_t slot = uint32_t((epoch - 1) % DistState::kRing);
  const size_t flag_bytes = win_flag_bytes(W);
  for (int r = 0; r < W; ++r) {
    if (r == D.rank || D.recv_blocks[r] == 0) continue;
    const uint32_t n = uint32_t(D.recv_blocks[r]);
    const double* in = reinterpret_cast<const double*>(D.win + flag_bytes + size_t(slot) * D.win_slot_bytes) +
                       size_t(D.recv_off[r]) * 36;
    apply_ghost_p2p_kernel<<<div_up(uint64_t(n) * 6, 128), 128, 0, h->stream>>>(
        n, D.recv_dst_block.p + D.recv_off[r], D.recv_full.p + D.recv_off[r], h->key_bits, h->blk_key.p, h->blk_off.p,
        h->node_len.p, h->node_base.p, in, h->values.p, win_arrived(D.win, r), epoch,
        win_consumed(D.peer_win[r], W, D.rank), D.done_count + W + r, p2p_timeout_ns(), 0x200u + uint32_t(r), D.d_err);
    h->launches++;
    D.last_recv += uint64_t(n) * 36 * 8;
  }
  FEMGPU_CUDA_CHECK(h, cudaGetLastError());
  return 0;
}

static int32_t p2p_numeric_exchange(Handle* h) {
  ++h->dist.epoch;
  int32_t st = p2p_pack(h, h->stream);
  if (st) return st;
  return p2p_apply(h);
}

// Ghost-first numeric pass (api.cu femgpu_numeric): the ghost rows are the LAST rows of the local matrix, hence the
// last slabs. When they are assembled first, their blocks can travel to the owners (pack kernel on a second stream)
// while the rest of the matrix is being assembled, and the owners' apply kernels find the flags already raised.
bool dist_ghost_first(Handle* h, uint32_t* first_ghost_slab) {
  DistState& D = h->dist;
  if (!D.enabled || !D.p2p || h->n_unstaged || h->n_ranges > 1 || h->n_slabs == 0) return false;
  static const bool off = getenv("FEMGPU_DIST_GHOST_FIRST") && atoi(getenv("FEMGPU_DIST_GHOST_FIRST")) == 0;
  if (off) return false;
  int64_t first_block = -1;
  for (int r = 0; r < D.world; ++r)
    if (r != D.rank && D.send_blocks[r] > 0 && (first_block < 0 || D.send_first_block[r] < first_block))
      first_block = D.send_first_block[r];
  if (first_block < 0) return false;  // nothing to send (the last rank of a strip partition)
  // slab k starts at the first node whose first block is >= k * quota, and the first ghost block is the first block
  // of the first ghost node: it lies in slab floor(first_block / quota)
  const uint32_t g0 = uint32_t(first_block / h->slab_quota);
  if (g0 == 0 || g0 >= h->n_slabs) return false;
  *first_ghost_slab = g0;
  return true;
}
int32_t dist_begin_pass(Handle* h) {
  h->dist.last_sent = h->dist.last_recv = 0;
  ++h->dist.epoch;
  return 0;
}
int32_t dist_pack(Handle* h, cudaStream_t st) { return p2p_pack(h, st); }
int32_t dist_apply(Handle* h) { return p2p_apply(h); }

static void p2p_close_peers(Handle* h) {
  for (auto& w : h->dist.peer_win) {
    if (w) cudaIpcCloseMemHandle(w);
    w = nullptr;
  }
  h->dist.p2p = false;
}
// An exported window may only be freed once no importer has it mapped any more: callers put a collective
// between p2p_close_peers() on every rank and this (dist_setup_p2p), or have synchronised the ranks themselves
// (femgpu_destroy: no rank may still be running numeric passes when a joined handle is destroyed).
static void p2p_free_window(unsigned char* win, uint32_t* done_count) {
  if (win) cudaFree(win);
  if (done_count) cudaFree(done_count);
}
static void p2p_release(Handle* h) {
  DistState& D = h->dist;
  p2p_close_peers(h);
  p2p_free_window(D.win, D.done_count);
  D.win = nullptr;
  D.done_count = nullptr;
}

// Collective (every rank calls it once the exchange plan of a symbolic pass is final): allocate and zero the
// window, trade IPC handles over the NCCL communicator, map the neighbours. All-or-nothing across the ranks.
int32_t dist_setup_p2p(Handle* h) {
  DistState& D = h->dist;
  const int W = D.world;
  // every rank has been through the collectives of this symbolic pass, i.e. its stream is past the last numeric
  // pass that touched the old windows: unmap the peers now, free the own old window after the next collective
  p2p_close_peers(h);
  unsigned char* old_win = D.win;
  uint32_t* old_count = D.done_count;
  D.win = nullptr;
  D.done_count = nullptr;
  D.epoch = 0;
  if (!D.h_err) {
    void* hp = nullptr;
    FEMGPU_CUDA_CHECK(h, cudaHostAlloc(&hp, 64, cudaHostAllocMapped));
    std::memset(hp, 0, 64);
    D.h_err = static_cast<volatile uint32_t*>(hp);
    FEMGPU_CUDA_CHECK(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&D.d_err), hp, 0));
  }
  *D.h_err = 0;
  bool want = true;
  if (const char* q = getenv("FEMGPU_DIST_P2P")) want = atoi(q) != 0;
  const int64_t n_recv = D.recv_off[W];
  auto slot_bytes = [](int64_t blocks) { return (size_t(blocks) * 36 * 8 + 255) & ~size_t(255); };
  D.win_slot_bytes = slot_bytes(n_recv);
  D.peer_slot_bytes.assign(W, 0);
  for (int r = 0; r < W; ++r) {  // what rank r receives in total: column r of the count matrix
    int64_t total = 0;
    for (int q = 0; q < W; ++q)
      if (q != r) total += D.count_matrix[size_t(q) * W + r];
    D.peer_slot_bytes[r] = slot_bytes(total);
  }
  D.win_bytes = win_flag_bytes(W) + DistState::kRing * D.win_slot_bytes;
  bool ok = want;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof mine);
  if (ok) {
    ok = cudaMalloc(reinterpret_cast<void**>(&D.win), D.win_bytes) == cudaSuccess &&
         cudaMalloc(reinterpret_cast<void**>(&D.done_count), size_t(2 * W) * 4) == cudaSuccess &&
         cudaMemsetAsync(D.win, 0, D.win_bytes, h->stream) == cudaSuccess &&
         cudaMemsetAsync(D.done_count, 0, size_t(2 * W) * 4, h->stream) == cudaSuccess &&
         cudaStreamSynchronize(h->stream) == cudaSuccess && cudaIpcGetMemHandle(&mine, D.win) == cudaSuccess;
    if (!ok) cudaGetLastError();
  }
  // [ok, pid, handle (8 words)] from everyone; the gather is also the barrier "all windows are zeroed"
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::vector<int64_t> send(10, 0), all(size_t(10) * W, 0);
  send[0] = ok ? 1 : 0;
  send[1] = int64_t(getpid());
  std::memcpy(&send[2], &mine, 64);
  int32_t st = dist_allgather_i64(h, send.data(), all.data(), 10);
  p2p_free_window(old_win, old_count);  // every importer unmapped it before entering the gather
  if (st) return st;
  for (int r = 0; r < W; ++r)
    if (!all[size_t(10) * r] || (r != D.rank && all[size_t(10) * r + 1] == int64_t(getpid()))) ok = false;
  D.peer_win.assign(W, nullptr);
  D.peer_recv_off.assign(W, 0);
  if (ok) {
    for (int r = 0; r < W && ok; ++r) {
      if (r == D.rank || (D.send_blocks[r] == 0 && D.recv_blocks[r] == 0)) continue;
      cudaIpcMemHandle_t hd;
      std::memcpy(&hd, &all[size_t(10) * r + 2], 64);
      void* mapped = nullptr;
      if (cudaIpcOpenMemHandle(&mapped, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        ok = false;
        break;
      }
      D.peer_win[r] = static_cast<unsigned char*>(mapped);
      // my run inside rank r's receive area starts after the runs of the lower source ranks
      int64_t off = 0;
      for (int q = 0; q < D.rank; ++q)
        if (q != r) off += D.count_matrix[size_t(q) * W + r];
      D.peer_recv_off[r] = off;
    }
  }
  // second round: did every rank map all its neighbours?
  int64_t mine_ok = ok ? 1 : 0;
  std::vector<int64_t> oks(W, 0);
  if ((st = dist_allgather_i64(h, &mine_ok, oks.data(), 1))) return st;
  for (int r = 0; r < W; ++r) ok = ok && oks[r] != 0;
  if (!ok) {
    p2p_release(h);  // NCCL send / recv exchange
    if (getenv("FEMGPU_DIST_INFO")) fprintf(stderr, "[femgpu dist] rank %d: peer-to-peer windows unavailable, using ncclSend/ncclRecv\n", D.rank);
    return 0;
  }
  D.p2p = true;
  if (getenv("FEMGPU_DIST_INFO"))
    fprintf(stderr, "[femgpu dist] rank %d/%d: p2p window %.2f MB (%lld blocks in, %lld out)\n", D.rank, W,
            double(D.win_bytes) / 1e6, (long long)n_recv, (long long)D.send_off[W]);
  return 0;
}

int32_t dist_check(Handle* h) {
  DistState& D = h->dist;
  if (!D.h_err) return 0;
  const uint32_t code = *D.h_err;
  if (!code) return 0;
  *D.h_err = 0;
  const int peer = int(code & 0xFFu);
  return h->fail(FEMGPU_ERR_NCCL,
                 std::string("ghost-row exchange timed out on rank ") + std::to_string(D.rank) + " in numeric pass " +
                     std::to_string(D.epoch) + ((code & 0x200u) ? ": the blocks of rank " : ": the ring slot of rank ") +
                     std::to_string(peer) + ((code & 0x200u) ? " never arrived" : " was never consumed") +
                     " (did every rank run the same number of femgpu_numeric passes?)");
}

int32_t dist_numeric_exchange(Handle* h) {
  DistState& D = h->dist;
  if (!D.enabled) return 0;
  const int W = D.world;
  D.last_sent = D.last_recv = 0;
  if (D.p2p) return p2p_numeric_exchange(h);
  ++D.epoch;
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
  p2p_release(h);
  if (h->dist.h_err) cudaFreeHost(const_cast<uint32_t*>(h->dist.h_err));
  h->dist.h_err = nullptr;
  h->dist.d_err = nullptr;
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
  {  // the ranks of one machine share its cores for their batched adds (torchrun exports LOCAL_WORLD_SIZE)
    int local = world;
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) {
      const int v = atoi(e);
      if (v > 0 && v <= world) local = v;
    }
    host_threads_share(unsigned(local));
  }
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

int32_t femgpu_dist_set_node_window(femgpu_t* h, uint32_t first_node_index) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (h->n_nodes() != 0) return h->fail(FEMGPU_ERR_USAGE, "femgpu_dist_set_node_window must be called before the first node is added");
  if (first_node_index > h->nodes_number) return h->fail(FEMGPU_ERR_USAGE, "node window starts beyond nodes_number");
  h->node_index_base = first_node_index;
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

int32_t femgpu_dist_info(femgpu_t* h, int32_t* p2p, uint64_t* passes) {
  if (!h) return FEMGPU_ERR_USAGE;
  if (p2p) *p2p = h->dist.p2p ? 1 : 0;
  if (passes) *passes = h->dist.epoch;
  return 0;
}

}  // extern "C"
