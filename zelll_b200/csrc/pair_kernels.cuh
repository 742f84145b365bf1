// pair_kernels.cuh -- half-shell pair enumeration and its consumers.
//
// Replaces GridCell::particle_pairs (iters.rs:238-241: intra_cell_pairs::<Half> ++
// inter_cell_pairs::<Half>) iterated over all non-empty cells (cellgrid.rs:338-340, and its rayon
// twin :447-451), fused with the consumers the reference benches put behind it:
//   * CountConsumer : `.filter(dsq <cmp> c2).count()`      (benches/cellgrid.rs:84-88)
//   * EmitConsumer  : the materialised (label, label) list (python/src/lib.rs:283-315)
//   * LjConsumer    : `.map(lj).sum()`                      (benches/lj.rs:42-47, 81-92)
//
// Half shell.  The reference's Half space is the first 13 of its 26 neighbour offsets, i.e. the
// cells with (dx,dy,dz) <lex (0,0,0) (flatindex.rs:55-65, iters.rs:58-63) -- "merely an
// implementation artifact" (iters.rs:110-112).  Any half space yields the same UNORDERED pair
// set; we use the z-major one, (dz,dy,dx) <lex (0,0,0), because with x-fastest cell ids its 13
// cells plus the home cell collapse into 5 runs of consecutive cells = 5 contiguous ranges of
// the cell-sorted record array:
//     A: (dz=-1, dy=-1, dx=-1..1)   B: (dz=-1, dy=0, dx=-1..1)   C: (dz=-1, dy=+1, dx=-1..1)
//     D: (dz= 0, dy=-1, dx=-1..1)   E: (dz= 0, dy=0, dx=-1..0)   <- ends with the home cell
//
// Work decomposition.  A CTA owns a tile of consecutive home cells.  When the tile's particles
// and its lower halo (all cells back to cell - (w0*w1 + w0 + 1)) fit the stage buffer -- always
// the case for the slab-shaped benchmark box -- they form ONE contiguous range of records,
// brought into shared memory by a single TMA bulk copy (cp.async.bulk, mbarrier completion).
// Otherwise (wide grids) the tile reads records through L1/L2 directly.  Inside a tile each warp
// takes home cells round-robin; its 32 lanes hold 32 candidate particles j of the 5 runs in
// registers while the home particles i are broadcast from shared memory, so every shared-memory
// read in the inner loop is a conflict-free broadcast.
//
// Arithmetic.  dsq = (dx*dx + dy*dy) + dz*dz with separately rounded operations, exactly what
// nalgebra::distance_squared does (benches/lj.rs:84); the TU is compiled with -fmad=false.
#pragma once

#include "common.cuh"

namespace zb {

constexpr int kPairThreads = 256;
constexpr int kPairWarps = kPairThreads / 32;
constexpr int kStageCells = 512;  // staged CSR entries per tile (cells + halo + 1)

template <class T>
struct PairParams {
  const Rec<T>* sorted;
  const uint32_t* csr;  // csr[c] .. csr[c+1] = records of cell c
  int w0, w1, w2;       // stored cells per axis
  uint32_t home_lo, home_hi;  // cell id range acting as home cells
  uint32_t tile_cells;
  uint32_t ntiles;
  uint32_t stage_recs;  // capacity of the record stage buffer
  T c2;                 // squared filter radius, in T (cutoff.powi(2))
};

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (PTX; SASS: SYNCS / UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}

// lj (benches/lj.rs:42-47): dsq.recip().powi(3) -> (r*r)*r ; 4*t*(t-1).  Exact IEEE division.
template <class T>
__device__ __forceinline__ T lj_term(T dsq) {
  T r = T(1) / dsq;
  T t = (r * r) * r;
  return (T(4) * t) * (t - T(1));
}

// ---------------------------------------------------------------------------------------------
// Consumers.  hit() is called by all 32 lanes of a warp in convergence, once per distance test.

struct ConsumerSmem {
  unsigned long long tile_count;  // CountConsumer
  uint32_t cursor;                // EmitConsumer
};

// -- count -----------------------------------------------------------------------------------
template <class T>
struct CountConsumer {
  struct Args {
    unsigned long long* tile_counts;  // [ntiles] or nullptr
    unsigned long long* block_totals; // [gridDim.x]
  };
  static constexpr int kWarpSmemBytes = 0;
  static constexpr int kStage = 4;  // ZB_STAGE_PAIR_COUNT
  Args a;
  ConsumerSmem* cs;
  unsigned long long cnt;    // per-lane, current tile
  unsigned long long total;  // thread 0: this block's running total

  __device__ CountConsumer(const Args& args, ConsumerSmem* s, void*) : a(args), cs(s), cnt(0), total(0) {}
  __device__ __forceinline__ void tile_begin(uint32_t) {
    if (threadIdx.x == 0) cs->tile_count = 0;
    cnt = 0;
  }
  __device__ __forceinline__ void hit(bool h, T, uint32_t, uint32_t) { cnt += h ? 1u : 0u; }
  __device__ __forceinline__ void tile_end(uint32_t tile) {
    unsigned long long w = warp_reduce(cnt, [](unsigned long long x, unsigned long long y) { return x + y; });
    if (lane_id() == 0 && w) atomicAdd(&cs->tile_count, w);
    __syncthreads();
    if (threadIdx.x == 0) {
      if (a.tile_counts) a.tile_counts[tile] = cs->tile_count;
      total += cs->tile_count;
    }
  }
  __device__ __forceinline__ void finish() {
    if (threadIdx.x == 0) a.block_totals[blockIdx.x] = total;
  }
};

// -- emit ------------------------------------------------------------------------------------
// Each tile owns the output range [tile_offsets[tile], tile_offsets[tile+1]) computed from a
// previous CountConsumer pass, so the list is compact and needs no global atomics: warps stage
// hits in a 64-entry shared queue and claim 32 slots at a time from the tile's shared cursor.
template <class T>
struct EmitConsumer {
  struct Args {
    const unsigned long long* tile_offsets;  // [ntiles + 1]
    uint2* out;
  };
  static constexpr int kWarpSmemBytes = 64 * sizeof(uint2);
  static constexpr int kStage = 5;  // ZB_STAGE_PAIR_EMIT
  Args a;
  ConsumerSmem* cs;
  uint2* q;
  int qn;
  unsigned long long base;

  __device__ EmitConsumer(const Args& args, ConsumerSmem* s, void* warp_smem)
      : a(args), cs(s), q(static_cast<uint2*>(warp_smem)), qn(0), base(0) {}
  __device__ __forceinline__ void tile_begin(uint32_t tile) {
    if (threadIdx.x == 0) cs->cursor = 0;
    base = a.tile_offsets[tile];
    qn = 0;
  }
  __device__ __forceinline__ void flush(int count) {
    __syncwarp();
    qn -= count;
    uint32_t pos = 0;
    if (lane_id() == 0) pos = atomicAdd(&cs->cursor, (uint32_t)count);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if ((int)lane_id() < count) a.out[base + pos + lane_id()] = q[qn + lane_id()];
    __syncwarp();
  }
  __device__ __forceinline__ void hit(bool h, T, uint32_t li, uint32_t lj) {
    unsigned b = __ballot_sync(0xffffffffu, h);
    if (b) {
      if (h) q[qn + __popc(b & lanemask_lt())] = make_uint2(li, lj);
      qn += __popc(b);
      if (qn >= 32) flush(32);
    }
  }
  __device__ __forceinline__ void tile_end(uint32_t) {
    if (qn > 0) flush(qn);
    __syncthreads();
  }
  __device__ __forceinline__ void finish() {}
};

// -- Lennard-Jones energy ----------------------------------------------------------------------
// Hits are ~20 % of the tests, so evaluating lj() under the hit predicate would run the
// division at ~6/32 lane efficiency.  Instead hits' dsq are compacted into a per-warp shared
// queue and evaluated 32 at a time.  Per-lane f64 partial sums -> warp shuffle -> block ->
// block_energy[blockIdx]; a last single-block kernel folds those in fixed order.
template <class T>
struct LjConsumer {
  struct Args {
    double* block_energy;               // [gridDim.x]
    unsigned long long* block_totals;   // [gridDim.x]
  };
  static constexpr int kWarpSmemBytes = 64 * sizeof(T);
  static constexpr int kStage = 6;  // ZB_STAGE_PAIR_LJ
  Args a;
  T* q;
  int qn;
  double acc;
  unsigned long long cnt;  // warp-uniform

  __device__ LjConsumer(const Args& args, ConsumerSmem*, void* warp_smem)
      : a(args), q(static_cast<T*>(warp_smem)), qn(0), acc(0.0), cnt(0) {}
  __device__ __forceinline__ void tile_begin(uint32_t) {}
  __device__ __forceinline__ void hit(bool h, T dsq, uint32_t, uint32_t) {
    unsigned b = __ballot_sync(0xffffffffu, h);
    if (b) {
      if (h) q[qn + __popc(b & lanemask_lt())] = dsq;
      int k = __popc(b);
      qn += k;
      cnt += (unsigned)k;
      if (qn >= 32) {
        __syncwarp();
        qn -= 32;
        T d = q[qn + lane_id()];
        acc += (double)lj_term(d);
        __syncwarp();
      }
    }
  }
  __device__ __forceinline__ void tile_end(uint32_t) { __syncthreads(); }
  __device__ __forceinline__ void finish() {
    __syncwarp();
    if ((int)lane_id() < qn) acc += (double)lj_term(q[lane_id()]);
    __shared__ double s_e[kPairWarps];
    __shared__ unsigned long long s_c[kPairWarps];
    double w = warp_reduce(acc, [](double x, double y) { return x + y; });
    if (lane_id() == 0) {
      s_e[threadIdx.x >> 5] = w;
      s_c[threadIdx.x >> 5] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double e = 0.0;
      unsigned long long c = 0;
      for (int i = 0; i < kPairWarps; ++i) { e += s_e[i]; c += s_c[i]; }
      a.block_energy[blockIdx.x] = e;
      a.block_totals[blockIdx.x] = c;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// One warp enumerates the half-shell pairs of home cell c.
//   STAGED: records and CSR entries come from shared memory (rec[-rec_origin], csr[-csr_origin]).
template <class T, int CMP, bool STAGED, class Consumer>
__device__ __forceinline__ void process_cell(const PairParams<T>& p, uint32_t c, const Rec<T>* __restrict__ rec,
                                             uint32_t rec_origin, const uint32_t* __restrict__ csr,
                                             uint32_t csr_origin, Consumer& cons) {
  auto CS = [&](uint32_t cell) -> uint32_t {
    return STAGED ? csr[cell - csr_origin] : __ldg(csr + cell);
  };
  const uint32_t hb = CS(c), he = CS(c + 1);
  const uint32_t m = he - hb;
  if (m == 0) return;

  const uint32_t w0 = (uint32_t)p.w0, w1 = (uint32_t)p.w1;
  const uint32_t cx = c % w0, r = c / w0;
  const uint32_t cy = r % w1, cz = r / w1;
  const uint32_t xl = cx > 0 ? 1u : 0u, xr = (cx + 1 < w0) ? 1u : 0u;
  const uint32_t plane = w0 * w1;

  // the 5 runs as record ranges [s, s+l)
  uint32_t sA = 0, lA = 0, sB = 0, lB = 0, sC = 0, lC = 0, sD = 0, lD = 0;
  if (cz > 0) {
    const uint32_t cb = c - plane;
    if (cy > 0) { sA = CS(cb - w0 - xl); lA = CS(cb - w0 + xr + 1) - sA; }
    { sB = CS(cb - xl); lB = CS(cb + xr + 1) - sB; }
    if (cy + 1 < w1) { sC = CS(cb + w0 - xl); lC = CS(cb + w0 + xr + 1) - sC; }
  }
  if (cy > 0) { sD = CS(c - w0 - xl); lD = CS(c - w0 + xr + 1) - sD; }
  const uint32_t sE = CS(c - xl), lE = he - sE;

  const uint32_t o1 = lA, o2 = o1 + lB, o3 = o2 + lC, o4 = o3 + lD, K = o4 + lE;
  const uint32_t shA = sA, shB = sB - o1, shC = sC - o2, shD = sD - o3, shE = sE - o4;
  const uint32_t first_home = K - m;  // candidates [first_home, K) are the home cell itself
  const unsigned lane = lane_id();
  const Rec<T>* home = rec + (hb - rec_origin);

  for (uint32_t kb = 0; kb < K; kb += 32) {
    const uint32_t k = kb + lane;
    const bool valid = k < K;
    const uint32_t sh = k < o1 ? shA : (k < o2 ? shB : (k < o3 ? shC : (k < o4 ? shD : shE)));
    const uint32_t pos = valid ? k + sh : hb;
    const Rec<T> rj = load_rec(rec + (pos - rec_origin));
    // position of candidate j inside the home cell; huge for other cells, 0 for idle lanes:
    // the pair (i, j) is taken iff u > i  (intra-cell: j after i, iters.rs:29-36)
    const uint32_t u = valid ? (k - first_home) : 0u;
#pragma unroll 2
    for (uint32_t i = 0; i < m; ++i) {
      const Rec<T> ri = load_rec(home + i);
      bool h = u > i;
      T dsq = T(0);
      if (CMP != 0) {
        const T dx = ri.x - rj.x, dy = ri.y - rj.y, dz = ri.z - rj.z;
        dsq = (dx * dx + dy * dy) + dz * dz;
        h = h && (CMP == 1 ? dsq < p.c2 : dsq <= p.c2);
      }
      cons.hit(h, dsq, ri.label, rj.label);
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <class T, int CMP, class Consumer>
__global__ void __launch_bounds__(kPairThreads) pair_kernel(PairParams<T> p, typename Consumer::Args args) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Rec<T>* s_rec = reinterpret_cast<Rec<T>*>(smem_raw);
  uint32_t* s_csr = reinterpret_cast<uint32_t*>(s_rec + p.stage_recs);
  unsigned char* s_cons = reinterpret_cast<unsigned char*>(s_csr + kStageCells + 4);
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ ConsumerSmem s_cs;

  const int warp = threadIdx.x >> 5;
  Consumer cons(args, &s_cs, s_cons + (size_t)warp * Consumer::kWarpSmemBytes);

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  uint32_t phase = 0;

  const uint32_t halo = (uint32_t)p.w0 * (uint32_t)p.w1 + (uint32_t)p.w0 + 1u;
  for (uint32_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
    const uint32_t c0 = p.home_lo + tile * p.tile_cells;
    const uint32_t c1 = min(c0 + p.tile_cells, p.home_hi);
    const uint32_t cl = c0 > halo ? c0 - halo : 0u;
    const uint32_t ncsr = c1 - cl + 1;
    const uint32_t plo = __ldg(p.csr + cl), phi = __ldg(p.csr + c1);
    const uint32_t np = phi - plo;
    const bool staged = np <= p.stage_recs && ncsr <= (uint32_t)kStageCells;

    cons.tile_begin(tile);
    if (staged) {
      if (threadIdx.x == 0 && np > 0) {
        const uint32_t bytes = np * (uint32_t)sizeof(Rec<T>);
        mbar_expect_tx(&s_bar, bytes);
        bulk_g2s(s_rec, p.sorted + plo, bytes, &s_bar);
      }
      for (uint32_t k = threadIdx.x; k < ncsr; k += kPairThreads) s_csr[k] = __ldg(p.csr + cl + k);
      __syncthreads();
      if (np > 0) {
        mbar_wait(&s_bar, phase);
        phase ^= 1u;
      }
      for (uint32_t c = c0 + warp; c < c1; c += kPairWarps)
        process_cell<T, CMP, true>(p, c, s_rec, plo, s_csr, cl, cons);
    } else {
      __syncthreads();
      for (uint32_t c = c0 + warp; c < c1; c += kPairWarps)
        process_cell<T, CMP, false>(p, c, p.sorted, 0u, p.csr, 0u, cons);
    }
    cons.tile_end(tile);  // ends with __syncthreads(): the stage buffers may be overwritten
  }
  cons.finish();
}

// ---------------------------------------------------------------------------------------------
// fixed-order fold of the per-block partials: out_e[0] = energy, out_c[0] = pair count
__global__ void finalize_kernel(const double* __restrict__ block_energy,
                                const unsigned long long* __restrict__ block_totals, uint32_t nblocks,
                                double* __restrict__ out_e, unsigned long long* __restrict__ out_c) {
  __shared__ double s_e[32];
  __shared__ unsigned long long s_c[32];
  double e = 0.0;
  unsigned long long c = 0;
  for (uint32_t b = threadIdx.x; b < nblocks; b += blockDim.x) {
    if (block_energy) e += block_energy[b];
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
    if (out_e) *out_e = te;
    if (out_c) *out_c = tc;
  }
}

// exclusive scan of the per-tile pair counts (a few 10^4 entries): one block, serial over chunks
__global__ void tile_offsets_kernel(const unsigned long long* __restrict__ counts, uint32_t ntiles,
                                    unsigned long long* __restrict__ offsets) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t b = 0; b < ntiles; b += blockDim.x) {
    uint32_t i = b + threadIdx.x;
    unsigned long long v = i < ntiles ? counts[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = lane < (int)(blockDim.x / 32) ? s_warp[lane] : 0ull;
      unsigned long long wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += y;
      }
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    unsigned long long carry = s_carry;
    if (i < ntiles) offsets[i] = carry + s_warp[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_warp[warp] + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[ntiles] = s_carry;
}

}  // namespace zb
