// p2p_kernels.cuh -- the three small exchanges of the slab step over NVLink peer memory, without NCCL.
//
// A slab step exchanges 6 doubles (global bounding box, max), one layer of halo
// rows with the neighbour, and 3 doubles (energy, pair count, error flag, sum).  With NCCL each of them is a
// launch of its own with tens of microseconds of latency at 8 GPUs -- more than the kernels around them leave
// idle.  Here every rank owns a MAILBOX in device memory that all peers of the node have mapped
// (cudaIpcGetMemHandle / cudaIpcOpenMemHandle at zb_comm_init) and the exchange is fused into the kernel that
// produces the data:
//
//   * all-reduce: the producer's last step writes its vector into slot [parity][rank] of EVERY peer's mailbox
//     (plain stores over NVLink), fences, and publishes a sequence number; it then waits until all `world`
//     slots of its own mailbox carry the current sequence number and folds them in rank order -- the same
//     order on every rank, so all ranks hold bit-identical results;
//   * halo: the rows a rank extracted while counting are stored straight into the neighbour's receive
//     block, followed by the sequence number; the neighbour's unpack kernel waits for it.
//
// Slots are double-buffered by the parity of the sequence number: a rank can be at most one exchange ahead of
// a peer (it needs that peer's contribution to finish the current one).  Waits are bounded (kP2pTimeoutNs):
// a dead peer raises bit 3 of the step's flag word instead of hanging the GPU.
#pragma once

#include "common.cuh"

namespace zb {

constexpr int kP2pMaxWorld = 16;
constexpr unsigned long long kP2pTimeoutNs = 20000000000ull;  // 20 s: ranks may reach their first exchange seconds apart

struct __align__(64) P2pSlot {
  double v[6];
  unsigned long long seq;
  unsigned long long pad;
};

// layout of one rank's mailbox allocation
struct P2pLayout {
  // [2][kP2pMaxWorld] box slots, [2][kP2pMaxWorld] energy slots, [2] halo flags (64 B apart), [2] halo blocks
  __host__ __device__ static constexpr size_t box_off(int parity, int r) { return ((size_t)parity * kP2pMaxWorld + r) * sizeof(P2pSlot); }
  __host__ __device__ static constexpr size_t energy_off(int parity, int r) { return box_off(2, 0) + box_off(parity, r); }
  __host__ __device__ static constexpr size_t halo_flag_off(int parity) { return 2 * box_off(2, 0) + (size_t)parity * 64; }
  __host__ __device__ static constexpr size_t halo_block_off(int parity, size_t block_bytes) { return 2 * box_off(2, 0) + 128 + (size_t)parity * block_bytes; }
  __host__ __device__ static constexpr size_t total(size_t block_bytes) { return 2 * box_off(2, 0) + 128 + 2 * block_bytes; }
};

struct P2pPeers {
  unsigned char* base[kP2pMaxWorld];  // mailbox of every rank as mapped into THIS process (own rank: local pointer)
};

__device__ __forceinline__ unsigned long long p2p_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *p == seq; false on timeout
__device__ __forceinline__ bool p2p_wait(const unsigned long long* p, unsigned long long seq) {
  const unsigned long long t0 = p2p_now();
  while (ld_acquire_sys(p) != seq) {
    if (p2p_now() - t0 > kP2pTimeoutNs) return false;
    __nanosleep(100);
  }
  return true;
}

// All-reduce of K <= 6 doubles over the node, called by ONE block at the end of the kernel that produced
// `mine` (threads 0 .. world-1 take part; `mine` must be visible to them: shared memory or registers of
// thread 0 broadcast by the caller).  MAXOP: max (bounding box as (-inf, sup)), else sum (energy, count, error
// flag).  Returns false on timeout (a peer never arrived).  out[] is written by thread 0.
template <int K, bool MAXOP>
__device__ __forceinline__ bool p2p_allreduce(const P2pPeers& peers, int world, int rank, unsigned long long seq, size_t slot0_off,
                                              const double* mine /* shared or global */, double* out) {
  __shared__ int s_ok;
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  const int r = threadIdx.x;
  const int parity = (int)(seq & 1ull);
  const size_t stride = sizeof(P2pSlot);
  if (r < world) {
    // publish my vector into slot [parity][rank] of peer r's mailbox
    P2pSlot* dst = reinterpret_cast<P2pSlot*>(peers.base[r] + slot0_off + ((size_t)parity * kP2pMaxWorld + rank) * stride);
#pragma unroll
    for (int k = 0; k < K; ++k) dst->v[k] = mine[k];
    __threadfence_system();
    st_release_sys(&dst->seq, seq);
    // wait for peer r's vector in my own mailbox
    const P2pSlot* src = reinterpret_cast<const P2pSlot*>(peers.base[rank] + slot0_off + ((size_t)parity * kP2pMaxWorld + r) * stride);
    if (!p2p_wait(&src->seq, seq)) atomicExch(&s_ok, 0);
  }
  __syncthreads();
  const bool ok = s_ok != 0;
  if (threadIdx.x == 0 && ok) {
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = MAXOP ? -INFINITY : 0.0;
    for (int q = 0; q < world; ++q) {  // rank order: the same fold on every rank
      const P2pSlot* src = reinterpret_cast<const P2pSlot*>(peers.base[rank] + slot0_off + ((size_t)parity * kP2pMaxWorld + q) * stride);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double v = src->v[k];
        acc[k] = MAXOP ? fmax(acc[k], v) : acc[k] + v;
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = acc[k];
  }
  return ok;
}

struct Box6v {
  double v[6];
};

// Global bounding box of a slab step in ONE launch behind K1: widen the local box to (-inf, sup) doubles,
// all-reduce(max) over peer memory, and -- speculative step -- compare with the box the step assumed
// (bit 2 of *flag = it changed).  box_t: K1's result in the grid's dtype; out6: the reduced box.
template <class T>
__global__ void p2p_box_kernel(P2pPeers peers, int world, int rank, unsigned long long seq, const T* __restrict__ box_t, int ndim,
                               int have_local, double* __restrict__ out6, int check, Box6v expect, uint32_t* __restrict__ flag) {
  __shared__ double s_mine[6];
  if (threadIdx.x < 6) {
    const int k = threadIdx.x;
    double v = -INFINITY;  // a rank without particles does not constrain the box
    if (have_local) {
      v = (k % 3) < ndim ? (double)box_t[k] : 0.0;
      if (k < 3) v = -v;
    }
    s_mine[k] = v;
  }
  __syncthreads();
  const bool ok = p2p_allreduce<6, true>(peers, world, rank, seq, P2pLayout::box_off(0, 0), s_mine, out6);
  if (threadIdx.x == 0) {
    if (!ok) {
      atomicOr(flag, 8u);
      for (int k = 0; k < 6; ++k) out6[k] = s_mine[k];
    } else if (check) {
      bool same = true;
      for (int k = 0; k < 6; ++k) same = same && (out6[k] == expect.v[k]);
      if (!same) atomicOr(flag, 4u);
    }
  }
}

// Energy of a slab step in ONE launch behind the LJ kernel: fold the per-block partials in fixed order
// (finalize_kernel's job), attach the step's own verdict, all-reduce(sum) over peer memory.
// out3 = (energy, pairs as f64, ranks that failed).  local_bad: the host already knows this rank failed.
__global__ void __launch_bounds__(256) p2p_energy_kernel(P2pPeers peers, int world, int rank, unsigned long long seq,
                                                         const double* __restrict__ block_energy,
                                                         const unsigned long long* __restrict__ block_totals, uint32_t nblocks,
                                                         int local_bad, const int* __restrict__ flags, uint32_t* __restrict__ slab_flag,
                                                         double* __restrict__ energy_out, unsigned long long* __restrict__ count_out,
                                                         double* __restrict__ out3) {
  __shared__ double s_e[8];
  __shared__ unsigned long long s_c[8];
  __shared__ double s_mine[3];
  double e = 0.0;
  unsigned long long c = 0;
  if (!local_bad)
    for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) {
      e += block_energy[b];
      c += block_totals[b];
    }
  e = warp_reduce(e, [](double x, double y) { return x + y; });
  c = warp_reduce(c, [](unsigned long long x, unsigned long long y) { return x + y; });
  if (lane_id() == 0) { s_e[threadIdx.x >> 5] = e; s_c[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double te = 0.0;
    unsigned long long tc = 0;
    for (unsigned w = 0; w < (blockDim.x + 31) / 32; ++w) { te += s_e[w]; tc += s_c[w]; }
    *energy_out = te;   // this rank's own share (zb_grid_lj_energy semantics)
    *count_out = tc;
    s_mine[0] = te;
    s_mine[1] = (double)tc;
    s_mine[2] = (local_bad || (*flags & 1) || (*slab_flag & 15u)) ? 1.0 : 0.0;
  }
  __syncthreads();
  const bool ok = p2p_allreduce<3, false>(peers, world, rank, seq, P2pLayout::energy_off(0, 0), s_mine, out3);
  if (threadIdx.x == 0 && !ok) {
    atomicOr(slab_flag, 8u);
    out3[0] = 0.0; out3[1] = 0.0; out3[2] = 1.0;
  }
}

// Halo push: the rows a rank extracted while counting -> the neighbour's receive block (row 0 = header with
// the row count, written here), then the sequence number.  Launched with a handful of blocks; the last one to
// finish publishes.
template <class T>
__global__ void __launch_bounds__(256) p2p_halo_push_kernel(const T* __restrict__ block, const uint32_t* __restrict__ count, uint32_t cap_rows,
                                                            T* __restrict__ peer_block, unsigned long long* __restrict__ peer_flag,
                                                            unsigned long long seq, unsigned* __restrict__ ticket) {
  const uint32_t cnt = *count;
  const uint32_t rows = min(cnt, cap_rows);  // an overflowing count is clamped; the header carries cap + 1 and the receiver flags it
  const uint32_t total = (rows + 1u) * 4u;
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < total; k += gridDim.x * blockDim.x)
    peer_block[k] = k == 0 ? (T)min(cnt, cap_rows + 1u) : (k < 4u ? T(0) : block[k]);
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1) {  // wraps back to 0: re-armed for the next step
      __threadfence_system();
      st_release_sys(peer_flag, seq);
    }
  }
}

// the receiver's side: wait until the neighbour has published this step's halo block
__global__ void p2p_halo_wait_kernel(const unsigned long long* __restrict__ flag_ptr, unsigned long long seq, uint32_t* __restrict__ flag) {
  if (threadIdx.x == 0 && !p2p_wait(flag_ptr, seq)) atomicOr(flag, 8u);
}

}  // namespace zb
