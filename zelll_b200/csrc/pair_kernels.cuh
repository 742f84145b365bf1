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
// Otherwise (wide grids) the tile reads records through L1/L2 directly.  The grid is persistent
// (SMs x occupancy CTAs): a CTA's first tile is blockIdx.x, every further one is claimed from a
// global counter one tile ahead.  Inside a tile one thread per home cell first computes the cell's
// run descriptor (CellRuns); warps then claim home cells dynamically.  A warp's 32 lanes hold up
// to 2 x 32 candidate particles j of the 5 runs in registers while the home particles i are
// broadcast from shared memory (explicit ld.shared through one 32-bit address register), two home
// particles per step; the last, partial chunk of a cell is packed (2 or 4 home particles tested
// against the same <= 16 / <= 8 candidates by different lane groups).
//
// Arithmetic.  dsq = (dx*dx + dy*dy) + dz*dz with separately rounded operations, exactly what
// nalgebra::distance_squared does (benches/lj.rs:84); the TU is compiled with -fmad=false.
//
// f64 grids can also run through the f32-prefiltered kernel of pair_pf_kernels.cuh (same enumeration, tests
// in packed f32 with a guard band, every "maybe" decided here-style in f64); work items that kernel declines
// come back to pair_kernel through PairParams::work_list.
//
// MODE splits the kernel so that the hot variant stays small (instruction cache): 1 = staged tiles only
// (tiles that do not fit the stage are appended to fb_list), 2 = global-memory tiles only (run over that
// list by a second launch), 0 = both in one kernel, 3 = compact cells of a sparse grid (sparse_kernels.cuh).
#pragma once

#include "common.cuh"

namespace zb {

#ifndef ZB_PAIR_THREADS
#define ZB_PAIR_THREADS 256
#endif
#ifndef ZB_PAIR_MINBLOCKS
#define ZB_PAIR_MINBLOCKS 3
#endif
#ifndef ZB_PAIR_UNROLL
#define ZB_PAIR_UNROLL 4
#endif
constexpr int kPairThreads = ZB_PAIR_THREADS;
constexpr int kPairUnroll = ZB_PAIR_UNROLL;
constexpr int kMaxNJ = 4;  // candidates held in registers per lane (register tile of the cell loop)
#ifndef ZB_LJ_FUSE
#define ZB_LJ_FUSE 2  // home particles per LJ compaction step (measured: 2 = 3 > 4 > 1)
#endif
constexpr int kPairWarps = kPairThreads / 32;
#ifndef ZB_TAIL_SPLIT
#define ZB_TAIL_SPLIT 16
#endif
constexpr uint32_t kTailSplit = ZB_TAIL_SPLIT;  // a remainder of up to this many candidates is packed (24 = split 17..24 into 16 + <= 8: measured neutral)
constexpr int kStageCells = 512;  // staged CSR entries per tile (cells + halo + 1)
constexpr int kMaxTileCells = 128;  // home cells per tile (their descriptors are staged)

// unsigned division by a launch-time constant (cell id -> cell coordinates) without the ~20
// instruction integer-division sequence: q = (t + ((n - t) >> sh1)) >> sh2, t = umulhi(n, mul)
struct FastDiv {
  uint32_t mul, sh1, sh2;
};
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(n, f.mul);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

template <class T>
struct PairParams {
  const Rec<T>* sorted;
  const uint32_t* csr;  // csr[c] .. csr[c+1] = records of cell c
  int w0, w1, w2;       // stored cells per axis
  uint32_t home_lo, home_hi;  // cell id range acting as home cells
  uint32_t tile_cells;
  uint32_t ntiles;
  // wide grids (a plane of cells holds far more records than the stage): tiles are segments of ONE x-row,
  // row_tiles of them per row (0 = tiles of consecutive cell ids across rows), and the stage holds the five
  // row segments the half shell needs instead of one contiguous range
  uint32_t row_tiles;
  uint32_t stage_recs;  // capacity of the record stage buffer
  T c2;                 // squared filter radius, in T (cutoff.powi(2))
  T fc;                 // filter radius
  T cell;               // edge length of a grid cell (the grid's cutoff)
  FastDiv div0, div1;   // division by w0 and by w1
  const uint32_t* tile_list;   // sparse boxes: ids of the tiles that hold home particles, else nullptr
  const uint32_t* tile_list_n; // ... and their number (device memory)
  uint32_t* tile_next;         // dynamic tile claim: work items handed out beyond the first wave; 0 between launches
  uint32_t* tile_done;         // CTAs that finished; the last one out re-arms both counters
  // prefilter kernel (pair_pf_kernels.cuh): work items it cannot take are appended here ...
  uint32_t* fb_list;
  uint32_t* fb_count;
  // ... and the exact kernel, launched behind it, walks exactly those (nullptr: all work items).  Per-tile
  // arrays stay indexed by the ORIGINAL work item.  The last CTA out clears the count for the next launch.
  const uint32_t* work_list;
  uint32_t* work_list_n;
  // sparse grids (sparse_kernels.cuh; MODE 3): "cells" are the compact non-empty cells u in [0, nuniq),
  // csr = ubegin, ukeys[u] = cx + w0 (cy + w1 cz) ascending; neighbour cells are found by binary search
  const unsigned long long* ukeys;
};

// home cells [c0, c1) of tile t
__device__ __forceinline__ void tile_cells_of(uint32_t t, uint32_t home_lo, uint32_t home_hi, uint32_t tile_cells,
                                              uint32_t row_tiles, uint32_t w0, uint32_t& c0, uint32_t& c1) {
  if (row_tiles) {
    const uint32_t row = t / row_tiles, seg = t - row * row_tiles;
    const uint32_t r0 = home_lo + row * w0;
    c0 = r0 + seg * tile_cells;
    c1 = min(c0 + tile_cells, r0 + w0);
  } else {
    c0 = home_lo + t * tile_cells;
    c1 = min(c0 + tile_cells, home_hi);
  }
}

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

// shared-memory access by 32-bit shared-space address (the hit queues of the consumers)
__device__ __forceinline__ void st_shared(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v));
}
__device__ __forceinline__ void st_shared(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v));
}
// a * b + c as ONE opaque instruction (keeps the compiler from re-associating the running queue
// address into separate index sums)
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
template <class T>
__device__ __forceinline__ T ld_shared(uint32_t a);
template <>
__device__ __forceinline__ double ld_shared<double>(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
template <>
__device__ __forceinline__ float ld_shared<float>(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}

// packed f32x2 arithmetic (sm_100: one FFMA2 = two IEEE fused multiply-adds, half the issue slots of two scalar
// instructions at the same pipe throughput).  fma(a, 1, b) and fma(a, a, 0) are the correctly rounded sum and
// product, so the reference's separately rounded f32 arithmetic can be expressed with them bit for bit.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpk2(uint64_t v, uint32_t& lo, uint32_t& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// lj (benches/lj.rs:42-47): dsq.recip().powi(3) -> (r*r)*r ; 4*t*(t-1).  Exact IEEE division.
template <class T>
__device__ __forceinline__ T lj_term(T dsq) {
  T r = T(1) / dsq;
  T t = (r * r) * r;
  return (T(4) * t) * (t - T(1));
}

// ---------------------------------------------------------------------------------------------
// small helpers

template <int N>
struct IntTag {
  static constexpr int value = N;
};

// candidates per lane of the exact loop.  Measured on B200 (n = 10^7 benchmark box): 2 per lane
// (one broadcast load of the home particle per 2 tests, ~78 registers, 3 CTAs/SM) is 3-5 % ahead
// of 1; 4 per lane needs ~128 registers and is slower.
#ifndef ZB_GENERIC_NJ
#define ZB_GENERIC_NJ 2
#endif
template <class T>
struct GenericNJ {
  static constexpr int value = ZB_GENERIC_NJ;
};

// opaque copy: keeps a kernel parameter in registers instead of re-reading the constant bank
__device__ __forceinline__ float keep_in_reg(float v) {
  asm volatile("" : "+f"(v));
  return v;
}
__device__ __forceinline__ double keep_in_reg(double v) {
  asm volatile("" : "+d"(v));
  return v;
}

// coordinates (and, when the consumer wants it, the label) of one record
template <bool LABEL>
__device__ __forceinline__ void load_part(const Rec<float>* p, float& x, float& y, float& z, uint32_t& label) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  x = v.x; y = v.y; z = v.z;
  label = LABEL ? __float_as_uint(v.w) : 0u;
}
template <bool LABEL>
__device__ __forceinline__ void load_part(const Rec<double>* p, double& x, double& y, double& z, uint32_t& label) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  x = a.x; y = a.y;
  if (LABEL) {
    const double2 b = reinterpret_cast<const double2*>(p)[1];
    z = b.x;
    label = (uint32_t)(__double_as_longlong(b.y) & 0xffffffffll);
  } else {
    z = reinterpret_cast<const double*>(p)[2];
    label = 0u;
  }
}

// Where the exact loop reads records from.  at(pos) rebases, load(pos) reads record `pos`.
//  * GlobalRecs: the cell-sorted array in global memory (tiles too large for the stage);
//  * StagedRecs: the shared-memory stage, addressed by a 32-bit shared-space BYTE address that is
//    biased by the stage's first record (wraps mod 2^32).  Explicit ld.shared keeps the address in
//    ONE register; through a generic pointer the compiler re-derived the shared window base
//    (S2UR SR_CgaCtaId + 3 uniform instructions) inside every test-loop iteration.
template <class T>
struct GlobalRecs {
  const Rec<T>* p;
  __device__ __forceinline__ GlobalRecs at(uint32_t pos) const { return GlobalRecs{p + pos}; }
  template <bool LABEL>
  __device__ __forceinline__ void load(uint32_t pos, T& x, T& y, T& z, uint32_t& label) const {
    load_part<LABEL>(p + pos, x, y, z, label);
  }
};
template <bool LABEL>
__device__ __forceinline__ void ld_shared_part(uint32_t a, float& x, float& y, float& z, uint32_t& label) {
  float w;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a));
  label = LABEL ? __float_as_uint(w) : 0u;
}
template <bool LABEL>
__device__ __forceinline__ void ld_shared_part(uint32_t a, double& x, double& y, double& z, uint32_t& label) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
  if (LABEL) {
    double w;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+16];" : "=d"(z), "=d"(w) : "r"(a));
    label = (uint32_t)__double2loint(w);
  } else {
    asm volatile("ld.shared.f64 %0, [%1+16];" : "=d"(z) : "r"(a));
    label = 0u;
  }
}
template <class T>
struct StagedRecs {
  uint32_t a;
  __device__ __forceinline__ StagedRecs at(uint32_t pos) const {
    return StagedRecs{a + pos * (uint32_t)sizeof(Rec<T>)};
  }
  template <bool LABEL>
  __device__ __forceinline__ void load(uint32_t pos, T& x, T& y, T& z, uint32_t& label) const {
    ld_shared_part<LABEL>(a + pos * (uint32_t)sizeof(Rec<T>), x, y, z, label);
  }
};

// the reference's distance_squared between two records, optionally with their labels
template <class T, bool LABEL>
__device__ __forceinline__ T exact_dsq(const Rec<T>* a, const Rec<T>* b, uint32_t& la, uint32_t& lb) {
  T xa, ya, za, xb, yb, zb_;
  load_part<LABEL>(a, xa, ya, za, la);
  load_part<LABEL>(b, xb, yb, zb_, lb);
  const T dx = xa - xb, dy = ya - yb, dz = za - zb_;
  return (dx * dx + dy * dy) + dz * dz;
}

template <int CMP, class T>
__device__ __forceinline__ bool passes(T dsq, T c2) {
  return CMP == 0 ? true : (CMP == 1 ? dsq < c2 : dsq <= c2);
}

// Guard band of the f32 prefilter.  Let o be the tile origin, R >= |x_p - o| for every staged
// record and axis, u = 2^-24.  The staged f32 coordinate r_p = fl32(x_p - o) is off by <= u R;
// d32 = fl32(r_i - r_j) is off the true difference d by <= 2 u R + u |d|; the three squares and two
// sums (FMA or not) add <= 3 u dsq.  Near the threshold |d| <= fc per axis and
// |dx| + |dy| + |dz| <= sqrt(3) fc, so |dsq32 - dsq| <= u (7 fc R + 6.5 fc^2): relative to fc^2 that
// is u (7 R / fc + 6.5).  We use delta = 2 u (8 R / fc + 8), more than twice the bound (this also
// covers the f64 roundings of the reference's own dsq and of x_p - o, ~2^-52, and the directed
// rounding of the two f32 thresholds).  dsq32 < c2 (1 - delta) => certainly inside; dsq32 >
// c2 (1 + delta) => certainly outside; in between the pair is decided in f64.
__device__ __forceinline__ float prefilter_delta(float R_over_fc) {
  return 1.1920929e-7f * (8.0f * R_over_fc + 8.0f);
}

// ---------------------------------------------------------------------------------------------
// Consumers.  test_n() is called by all 32 lanes of a warp in convergence, once per (fused) home particle
// step with the candidates a lane holds; hit() is the hand-over point of pf_pair_kernel.

// EmitConsumer's output description (explained there)
#ifndef ZB_EMIT_CHUNK
#define ZB_EMIT_CHUNK 512
#endif
constexpr uint32_t kEmitChunk = ZB_EMIT_CHUNK;  // rows per output slot (a multiple of 32)
constexpr uint32_t kEmitNoSlot = 0xffffffffu;

struct EmitOut {
  uint2* out;
  uint2* spill;
  uint32_t out_chunks, spill_chunks;
  unsigned long long* cursor;  // slots handed out
  uint2* partial;              // (slot, rows used) of partly filled slots
  uint32_t* npartial;
  uint32_t partial_cap;
  __device__ __forceinline__ uint2* rows(uint32_t slot) const {
    if (slot < out_chunks) return out + (size_t)slot * kEmitChunk;
    slot -= out_chunks;
    return slot < spill_chunks ? spill + (size_t)slot * kEmitChunk : nullptr;
  }
};

struct ConsumerSmem {
  unsigned long long tile_count;  // CountConsumer
  EmitOut emit;                   // EmitConsumer: the fields of its Args that only the rare paths need
};

// -- count -----------------------------------------------------------------------------------
template <class T>
struct CountConsumer {
  struct Args {
    unsigned long long* tile_counts;  // [ntiles] or nullptr
    unsigned long long* block_totals; // [gridDim.x]
  };
  static constexpr int kWarpSmemBytes = 0;
  static constexpr int kPfWarpSmemBytes = 0;  // per-warp bytes under pf_pair_kernel
  static constexpr int kStage = 4;  // ZB_STAGE_PAIR_COUNT
  static constexpr bool kNeedLabels = false;
  static constexpr bool kCountsOnly = true;
  static constexpr int kUnroll = kPairUnroll;  // test-loop unroll factor
  static constexpr int kMinBlocks = 4;         // CTAs per SM (64 registers): count is 3 % faster at 4, LJ / emit at 3
  static constexpr int kFuse = 1;              // home particles handed to test_n per call
  Args a;
  ConsumerSmem* cs;
  uint32_t c32;              // per-lane, current chunk (one predicated add per test)
  unsigned long long cnt;    // per-lane, current tile
  unsigned long long total;  // thread 0: this block's running total

  __device__ CountConsumer(const Args& args, ConsumerSmem* s, void*, T) : a(args), cs(s), c32(0), cnt(0), total(0) {}
  __device__ __forceinline__ void tile_begin(uint32_t, const Rec<T>*, bool) {
    if (threadIdx.x == 0) cs->tile_count = 0;
    cnt = 0;
  }
  template <int NJ>
  __device__ __forceinline__ void test_n(const bool (&h)[NJ], const T (&)[NJ], const uint32_t (&)[NJ],
                                         const uint32_t (&)[NJ]) {
    // one predicated add per test (plain C++ compiles to an add plus a predicated move)
#pragma unroll
    for (int q = 0; q < NJ; ++q)
      asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(c32) : "r"((uint32_t)h[q]));
  }
  __device__ __forceinline__ void add(uint32_t k) { cnt += k; }  // unfiltered candidates, no loop
  __device__ __forceinline__ void chunk_end() {
    cnt += c32;
    c32 = 0;
  }
  template <int CMP>
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
// ONE pass, no sizing pass in front of it.  The output is handed out in SLOTS of kEmitChunk rows from one
// global cursor, one slot per WARP at a time: a warp writes its 32-row stores into its own slot without any
// atomic, and fetches the slot after ahead of need (the atomic's result is first looked at when the current
// slot is full, thousands of cycles later).  One global atomic per kEmitChunk rows -- a same-address atomic per
// 32-row store (5e6 of them at n = 1e7) would serialise in L2, and slots shared by the warps of a CTA cost
// ~0.9 us of CTA time per slot switch (measured: 1.85 ms at 1024-row slots, 2.8 ms at 256).
// Slots [0, out_chunks) are the caller's buffer, the next spill_chunks a scratch area (a pass needs up to two
// slots per warp more than the list is long).  Every warp ends with at most two partly filled slots (its
// current one and the one it fetched ahead); they are recorded in `partial` and the emit_fix_* kernels below
// close the holes (order is unspecified upstream, iters.rs:251, so rows may move): full slots from behind the
// end of the list move into the partial slots in front of it, the partial rows are packed behind the last
// full slot.  A list longer than the capacity keeps counting slots without writing, so the caller still
// learns the size it needs.
__device__ __noinline__ uint2* emit_spill_rows(const ConsumerSmem* cs, uint32_t slot) {
  slot -= cs->emit.out_chunks;
  return slot < cs->emit.spill_chunks ? cs->emit.spill + (size_t)slot * kEmitChunk : nullptr;
}

template <class T>
struct EmitConsumer {
  using Args = EmitOut;
  static constexpr int kWarpSmemBytes = (32 + 32 * kMaxNJ * 2) * sizeof(uint2);  // one row + one fused step
  static constexpr int kPfWarpSmemBytes = 64 * sizeof(uint2);  // pf_pair_kernel hands over one row at a time
  static constexpr int kStage = 5;  // ZB_STAGE_PAIR_EMIT
  static constexpr bool kNeedLabels = true;
  static constexpr bool kCountsOnly = false;
  static constexpr int kUnroll = 2;
  static constexpr int kFuse = 2;
  static constexpr int kMinBlocks = ZB_PAIR_MINBLOCKS;
  // only what the 32-row store needs lives in registers; the rest of Args waits in shared memory (cs->emit)
  uint2* out;
  uint32_t out_chunks;
  ConsumerSmem* cs;
  uint2* q;        // (label, label) rows waiting for a full 32-row store; survives tile changes
  uint32_t qn;
  unsigned ltmask;
  uint32_t slot, used;  // this warp's output slot and the rows of it that are written
  uint32_t next_l0;     // lane 0: the slot fetched ahead
  bool has_next;

  __device__ static __forceinline__ uint32_t fetch_slot(unsigned long long* cursor) {
    uint32_t v = 0;
    if (lane_id() == 0) v = (uint32_t)atomicAdd(cursor, 1ull);
    return v;
  }
  __device__ EmitConsumer(const Args& args, ConsumerSmem* s, void* warp_smem, T)
      : out(args.out), out_chunks(args.out_chunks), cs(s), q(static_cast<uint2*>(warp_smem)), qn(0), ltmask(lanemask_lt()),
        slot(kEmitNoSlot), used(kEmitChunk), next_l0(fetch_slot(args.cursor)), has_next(true) {
    if (threadIdx.x == 0) cs->emit = args;  // ordered before its first use by the barriers of the tile set-up
  }
  __device__ __forceinline__ void tile_begin(uint32_t, const Rec<T>*, bool) {}
  // destination of this warp's next 32 rows (nullptr: beyond the capacity)
  __device__ __forceinline__ uint2* claim32() {
    if (used == kEmitChunk) {  // warp-uniform
      slot = __shfl_sync(0xffffffffu, next_l0, 0);
      used = 0;
      next_l0 = fetch_slot(cs->emit.cursor);
    }
    uint2* base = slot < out_chunks ? out + (size_t)slot * kEmitChunk : emit_spill_rows(cs, slot);
    uint2* dst = base ? base + used : nullptr;
    used += 32;
    return dst;
  }
  __device__ __forceinline__ void drain_exact_rows() {
    while (qn >= 32) {
      __syncwarp();
      qn -= 32;
      const uint2 row = q[qn + lane_id()];
      uint2* dst = claim32();
      if (dst) dst[lane_id()] = row;
      __syncwarp();
    }
  }
  template <int NJ>
  __device__ __forceinline__ void test_n(const bool (&h)[NJ], const T (&)[NJ], const uint32_t (&li)[NJ],
                                         const uint32_t (&lj)[NJ]) {
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const unsigned b = __ballot_sync(0xffffffffu, h[k]);
      if (h[k]) q[qn + __popc(b & ltmask)] = make_uint2(li[k], lj[k]);
      qn += __popc(b);
    }
    drain_exact_rows();
  }
  // pair_pf_kernels.cuh: one exactly decided pair per lane
  __device__ __forceinline__ void hit(bool h, T, uint32_t li, uint32_t lj) {
    const unsigned b = __ballot_sync(0xffffffffu, h);
    if (h) q[qn + __popc(b & ltmask)] = make_uint2(li, lj);
    qn += __popc(b);
    drain_exact_rows();
  }
  __device__ __forceinline__ void add(uint32_t) {}
  __device__ __forceinline__ void chunk_end() {}
  template <int CMP>
  __device__ __forceinline__ void tile_end(uint32_t) {
    __syncthreads();
  }
  // the < 32 rows the warp is left with go behind its last store; what stays unused of its slots is recorded
  __device__ __noinline__ void finish() {
    const EmitOut& a = cs->emit;
    auto note_partial = [&](uint32_t sl, uint32_t rows) {
      const uint32_t i = atomicAdd(a.npartial, 1u);
      if (i < a.partial_cap) a.partial[i] = make_uint2(sl, rows);
    };
    __syncwarp();
    uint32_t nx = __shfl_sync(0xffffffffu, next_l0, 0);
    if (qn > 0 && used == kEmitChunk) {
      slot = nx;
      used = 0;
      has_next = false;
    }
    if (qn > 0) {
      uint2* base = a.rows(slot);
      if (base && lane_id() < qn) base[used + lane_id()] = q[lane_id()];
      used += qn;
    }
    if (lane_id() == 0) {
      if (slot != kEmitNoSlot && used < kEmitChunk) note_partial(slot, used);
      if (has_next) note_partial(nx, 0u);
    }
  }
};

// what emit_fix_plan_kernel decides
struct EmitFix {
  unsigned long long total;   // rows of the list
  uint32_t full;              // F: slots [0, F) are full after the moves, the partial rows follow at F * kEmitChunk
  uint32_t moves;             // full slots behind F that move into partial slots in front of it
  uint32_t npartial;
  uint32_t ok;                // 0: longer than the capacity (nothing is moved) or internal table overflow (2)
  unsigned long long tail_rows;  // rows of all partial slots
};

// One block.  partial[] = (slot, used) in arbitrary order; writes the packed offset of every partial slot
// (poff[i], rows), the (source, destination) slot of every move, and the verdict.
__global__ void __launch_bounds__(1024) emit_fix_plan_kernel(EmitOut o, unsigned long long cap_rows, uint32_t* __restrict__ poff,
                                                             uint2* __restrict__ moves, EmitFix* __restrict__ fix) {
  extern __shared__ unsigned char s_tailflag[];  // [partial_cap]: 1 = the slot F + t is a partial one
  __shared__ uint32_t s_scan[32];
  __shared__ uint32_t s_carry, s_nsrc, s_ndst;
  const uint32_t noted = *o.npartial;
  const uint32_t P = min(noted, o.partial_cap);
  const unsigned long long R = *o.cursor;
  const unsigned long long F = R - P;
  if (threadIdx.x == 0) { s_carry = 0; s_nsrc = 0; s_ndst = 0; }
  for (uint32_t t = threadIdx.x; t < P; t += blockDim.x) s_tailflag[t] = 0;
  __syncthreads();
  // packed offsets: exclusive scan of `used` in list order
  for (uint32_t base = 0; base < P; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < P ? o.partial[i].y : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, incl, d);
      if ((int)lane_id() >= d) incl += up;
    }
    if (lane_id() == 31) s_scan[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = s_scan[threadIdx.x], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t up = __shfl_up_sync(0xffffffffu, wi, d);
        if ((int)threadIdx.x >= d) wi += up;
      }
      s_scan[threadIdx.x] = wi - w;
    }
    __syncthreads();
    const uint32_t carry = s_carry;
    if (i < P) {
      poff[i] = carry + s_scan[threadIdx.x >> 5] + incl - v;
      const uint32_t slot = o.partial[i].x;
      if (slot >= F) s_tailflag[slot - F] = 1;                       // slot < R = F + P
      else moves[atomicAdd(&s_ndst, 1u)].y = slot;                   // a hole in front of F
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + s_scan[threadIdx.x >> 5] + incl;
    __syncthreads();
  }
  // the full slots behind F fill the holes in front of it (as many of the one as of the other)
  for (uint32_t t = threadIdx.x; t < P; t += blockDim.x)
    if (!s_tailflag[t]) moves[atomicAdd(&s_nsrc, 1u)].x = (uint32_t)(F + t);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long total = F * kEmitChunk + s_carry;
    fix->total = total;
    fix->full = (uint32_t)F;
    fix->moves = s_nsrc;
    fix->npartial = P;
    fix->tail_rows = s_carry;
    uint32_t ok = 1;
    if (total > cap_rows || R > (unsigned long long)o.out_chunks + o.spill_chunks) ok = 0;
    if (noted > o.partial_cap || s_nsrc != s_ndst) ok = 2;
    fix->ok = ok;
  }
}

// rows of every partial slot -> tmp (packed)
__global__ void __launch_bounds__(256) emit_fix_save_kernel(EmitOut o, const uint32_t* __restrict__ poff, const EmitFix* __restrict__ fix,
                                                            uint2* __restrict__ tmp) {
  if (fix->ok != 1) return;
  for (uint32_t i = blockIdx.x; i < fix->npartial; i += gridDim.x) {
    const uint2 pu = o.partial[i];
    const uint2* src = o.rows(pu.x);
    for (uint32_t k = threadIdx.x; k < pu.y; k += blockDim.x) tmp[poff[i] + k] = src[k];
  }
}
// full slots from behind the end of the list -> the partial slots in front of it
__global__ void __launch_bounds__(256) emit_fix_move_kernel(EmitOut o, const uint2* __restrict__ moves, const EmitFix* __restrict__ fix) {
  if (fix->ok != 1) return;
  for (uint32_t m = blockIdx.x; m < fix->moves; m += gridDim.x) {
    const uint2* src = o.rows(moves[m].x);
    uint2* dst = o.rows(moves[m].y);
    for (uint32_t k = threadIdx.x; k < kEmitChunk; k += blockDim.x) dst[k] = src[k];
  }
}
// tmp -> behind the last full slot
__global__ void __launch_bounds__(256) emit_fix_pack_kernel(EmitOut o, const uint2* __restrict__ tmp, const EmitFix* __restrict__ fix) {
  if (fix->ok != 1) return;
  uint2* dst = o.out + (size_t)fix->full * kEmitChunk;
  for (unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; k < fix->tail_rows;
       k += (unsigned long long)gridDim.x * blockDim.x)
    dst[k] = tmp[k];
}

// -- Lennard-Jones energy ----------------------------------------------------------------------
// Hits are ~20 % of the tests, so evaluating lj() under the hit predicate would run the
// division at ~6/32 lane efficiency.  Instead hits are compacted (ballot + popc) into a per-warp
// shared queue and evaluated 32 at a time with all lanes busy: the exact loop queues dsq, the
// prefilter loop queues the pair's tile-local positions and the drain recomputes dsq in f64.
// Per-lane f64 partial sums -> warp shuffle -> block -> block_energy[blockIdx]; finalize_kernel
// folds those in fixed order.
template <class T>
struct LjConsumer {
  struct Args {
    double* block_energy;               // [gridDim.x]
    unsigned long long* block_totals;   // [gridDim.x]
  };
  // (32 rows + one fused step) dsq values per warp
#ifndef ZB_LJ_DRAIN_ROWS
#define ZB_LJ_DRAIN_ROWS 2
#endif
  static constexpr int kDrainRows = ZB_LJ_DRAIN_ROWS;
  static constexpr int kExactBytes = (32 * kDrainRows + 32 * GenericNJ<T>::value * ZB_LJ_FUSE) * (int)sizeof(T);
  static constexpr int kWarpSmemBytes = kExactBytes;
  static constexpr int kPfWarpSmemBytes = 16;  // pf_pair_kernel evaluates rows in place (no dsq queue)
  static constexpr int kStage = 6;  // ZB_STAGE_PAIR_LJ
  static constexpr bool kNeedLabels = false;
  static constexpr bool kCountsOnly = false;
#ifndef ZB_LJ_UNROLL
#define ZB_LJ_UNROLL 2
#endif
  static constexpr int kUnroll = ZB_LJ_UNROLL;  // keeps the kernel below ~5k SASS instructions (I-cache)
  static constexpr int kMinBlocks = ZB_PAIR_MINBLOCKS;
  // 2 home particles per compaction step: their distance tests are independent instruction chains
  // that overlap, and the queue bookkeeping / drain check is paid once for both
  static constexpr int kFuse = ZB_LJ_FUSE;
  Args a;
  uint32_t q0;   // shared-space byte address of this warp's queue of dsq values ...
  uint32_t qa;   // ... and of its fill position (warp-uniform)
  unsigned ltmask;
  double acc;
  double acc4;             // pf_pair_kernel: sum of lj / 4
  unsigned long long cnt;  // per-lane pairs kept

  __device__ LjConsumer(const Args& args, ConsumerSmem*, void* warp_smem, T c2)
      : a(args), q0(smem_u32(warp_smem)), qa(q0), ltmask(lanemask_lt()), acc(0.0), acc4(0.0), cnt(0) {}
  // what the exact loop left in the queue (< kDrainRows rows of dsq values)
  __device__ __forceinline__ void drain_exact_leftovers() {
    __syncwarp();
    const uint32_t left = (qa - q0) / (uint32_t)sizeof(T);
    for (uint32_t r = 0; r < left; r += 32) {
      if (r + lane_id() < left) {
        acc += (double)lj_term(ld_shared<T>(q0 + (r + lane_id()) * (uint32_t)sizeof(T)));
        cnt += 1;
      }
    }
    qa = q0;
    __syncwarp();
  }
  __device__ __forceinline__ void tile_begin(uint32_t, const Rec<T>*, bool) {}
  template <int NJ>
  __device__ __forceinline__ void test_n(const bool (&h)[NJ], const T (&dsq)[NJ], const uint32_t (&)[NJ],
                                         const uint32_t (&)[NJ]) {
    // the fill position is kept as a shared-space BYTE address: slot address = one multiply-add,
    // advancing it = one multiply-add (an element index costs an extra add per test)
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      const unsigned b = __ballot_sync(0xffffffffu, h[k]);
      if (h[k]) st_shared(mad_u32(__popc(b & ltmask), sizeof(T), qa), dsq[k]);
      qa = mad_u32(__popc(b), sizeof(T), qa);
    }
    // drain two rows at a time: their reciprocal / power chains are independent and overlap.  One
    // copy of the lj code (the kernel must stay small).
    constexpr uint32_t kRowsBytes = kDrainRows * 32 * sizeof(T);
#pragma unroll 1
    while (qa >= q0 + kRowsBytes) {
      __syncwarp();
      qa -= kRowsBytes;
      T e[kDrainRows];
#pragma unroll
      for (int r = 0; r < kDrainRows; ++r)
        e[r] = lj_term(ld_shared<T>(qa + (32 * r + lane_id()) * (uint32_t)sizeof(T)));
#pragma unroll
      for (int r = 0; r < kDrainRows; ++r) acc += (double)e[r];
      cnt += kDrainRows;
      __syncwarp();
    }
  }
  // pair_pf_kernels.cuh: one exactly decided pair per lane (h = it passed the filter), evaluated in place.
  // acc4 collects t (t - 1) = lj / 4 with one fused multiply-add per pair; finish() scales by 4 (exact).
  // The reciprocal is two Newton steps on the hardware seed (relative error ~1e-16 per pair; the energy
  // tolerance of the north star is 1e-10 on the sum).
  __device__ __forceinline__ void hit(bool h, T dsq, uint32_t, uint32_t) {
    const double d = h ? (double)dsq : 1.0;  // idle lanes: a harmless finite value
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-d, r, 1.0);
    r = __fma_rn(r, e, r);
    const double t = (r * r) * r;
    if (h) {
      acc4 = __fma_rn(t, t - 1.0, acc4);
      cnt += 1;
    }
  }
  __device__ __forceinline__ void add(uint32_t) {}
  __device__ __forceinline__ void chunk_end() {}
  template <int CMP>
  __device__ __forceinline__ void tile_end(uint32_t) {
    // (the queue holds dsq values, which stay valid across tiles: they wait for finish())
    __syncthreads();
  }
  __device__ __forceinline__ void finish() {
    __shared__ double s_e[kPairWarps];
    __shared__ unsigned long long s_c[kPairWarps];
    drain_exact_leftovers();
    acc += 4.0 * acc4;
    const double w = warp_reduce(acc, [](double x, double y) { return x + y; });
    const unsigned long long c = warp_reduce(cnt, [](unsigned long long x, unsigned long long y) { return x + y; });
    if (lane_id() == 0) {
      s_e[threadIdx.x >> 5] = w;
      s_c[threadIdx.x >> 5] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double e = 0.0;
      unsigned long long t = 0;
      for (int i = 0; i < kPairWarps; ++i) { e += s_e[i]; t += s_c[i]; }
      a.block_energy[blockIdx.x] = e;
      a.block_totals[blockIdx.x] = t;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// The candidate set of home cell c: 5 runs of consecutive records (see the header comment).
// The descriptor is computed ONCE per cell by one thread at tile start (cell_runs) and kept in
// shared memory; the warp that later claims the cell reads it with three broadcast LDS.128 instead
// of redoing ~130 instructions of index arithmetic in all 32 lanes.
struct __align__(16) CellRuns {
  uint32_t hb, m;                 // home cell: first record, size
  uint32_t K, shE;                // number of candidates; candidates [K - m, K) are the home cell itself
  uint32_t o1, o2, o3, o4;        // run boundaries in candidate numbering
  uint32_t shA, shB, shC, shD;    // candidate k of run X is record k + shX

  // record of candidate k, or hb for an idle lane
  __device__ __forceinline__ uint32_t pos(uint32_t k) const {
    const uint32_t sh = k < o1 ? shA : (k < o2 ? shB : (k < o3 ? shC : (k < o4 ? shD : shE)));
    return k < K ? k + sh : hb;
  }
  // number of leading home particles candidate k pairs with: all m for a candidate of another
  // cell, the u particles before it for the home cell's u-th particle (intra-cell pairs: j
  // after i, iters.rs:29-36), none for an idle lane
  __device__ __forceinline__ uint32_t thr(uint32_t k) const {
    const uint32_t first_home = K - m;
    return k < K ? (k >= first_home ? k - first_home : m) : 0u;
  }
};

template <class T>
__device__ __forceinline__ bool cell_runs(const PairParams<T>& p, uint32_t c, const uint32_t* __restrict__ csrb,
                                          CellRuns& r) {
  const uint32_t hb = csrb[c], he = csrb[c + 1];
  r.hb = hb;
  r.m = he - hb;
  r.K = r.o1 = r.o2 = r.o3 = r.o4 = 0;
  r.shA = r.shB = r.shC = r.shD = r.shE = 0;
  if (r.m == 0) return false;
  const uint32_t w0 = (uint32_t)p.w0, w1 = (uint32_t)p.w1;
  const uint32_t row = fast_div(c, p.div0), cx = c - row * w0;
  const uint32_t cz = fast_div(row, p.div1), cy = row - cz * w1;
  const uint32_t xl = cx > 0 ? 1u : 0u, xr = (cx + 1 < w0) ? 1u : 0u;
  uint32_t sA = 0, lA = 0, sB = 0, lB = 0, sC = 0, lC = 0, sD = 0, lD = 0;
  if (cz > 0) {
    const uint32_t* cb = csrb + (c - w0 * w1);
    if (cy > 0) { sA = cb[-(int)(w0 + xl)]; lA = cb[-(int)w0 + (int)xr + 1] - sA; }
    { sB = cb[-(int)xl]; lB = cb[xr + 1] - sB; }
    if (cy + 1 < w1) { sC = cb[w0 - xl]; lC = cb[w0 + xr + 1] - sC; }
  }
  if (cy > 0) { sD = csrb[c - w0 - xl]; lD = csrb[c - w0 + xr + 1] - sD; }
  const uint32_t sE = xl ? csrb[c - 1] : hb, lE = he - sE;
  r.o1 = lA; r.o2 = r.o1 + lB; r.o3 = r.o2 + lC; r.o4 = r.o3 + lD; r.K = r.o4 + lE;
  r.shA = sA; r.shB = sB - r.o1; r.shC = sC - r.o2; r.shD = sD - r.o3; r.shE = sE - r.o4;
  return true;
}

// the same descriptor for compact cell u of a sparse grid: each run is the record range of (up to) three
// consecutive keys, found by binary search among the cells before u (all half-shell cells have smaller keys)
__device__ __forceinline__ uint32_t lower_bound_ukeys(const unsigned long long* __restrict__ a, uint32_t hi,
                                                      unsigned long long key) {
  uint32_t lo = 0;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(a + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}
template <class T>
__device__ __forceinline__ bool cell_runs_sparse(const PairParams<T>& p, uint32_t u, CellRuns& r) {
  const uint32_t hb = __ldg(p.csr + u), he = __ldg(p.csr + u + 1);
  r.hb = hb;
  r.m = he - hb;
  r.K = r.o1 = r.o2 = r.o3 = r.o4 = 0;
  r.shA = r.shB = r.shC = r.shD = r.shE = 0;
  if (r.m == 0) return false;
  const unsigned long long w0 = (unsigned long long)p.w0, w1 = (unsigned long long)p.w1;
  const unsigned long long key = __ldg(p.ukeys + u);
  const unsigned long long cx = key % w0, row = key / w0, cy = row % w1, cz = row / w1;
  const unsigned long long xl = cx > 0 ? 1u : 0u, xr = (cx + 1 < w0) ? 1u : 0u;
  // records of the cells with keys [k0, k1] among the compact cells before `hi` (<= u).  The wanted cells
  // sit a row or a plane of NON-EMPTY cells before u, usually far fewer than u itself: gallop backwards from
  // `hi` to bracket k0, then bisect inside the bracket.  Returns the first compact cell >= k0 so that the
  // next (smaller-keyed) run can start its search there.
  auto range = [&](uint32_t hi, unsigned long long k0, unsigned long long k1, uint32_t& s, uint32_t& l) -> uint32_t {
    uint32_t lo = hi, step = 1;
    while (lo > 0) {
      const uint32_t probe = lo > step ? lo - step : 0u;
      if (__ldg(p.ukeys + probe) < k0) {  // bracket found: first cell >= k0 lies in (probe, lo]
        const uint32_t first = probe + 1 + lower_bound_ukeys(p.ukeys + probe + 1, lo - probe - 1, k0);
        lo = first;
        goto found;
      }
      lo = probe;
      step <<= 1;
    }
    lo = 0;  // every cell before hi has a key >= k0
  found:
    uint32_t up = lo;
    while (up < hi && up < lo + 3u && __ldg(p.ukeys + up) <= k1) ++up;
    s = __ldg(p.csr + lo);
    l = __ldg(p.csr + up) - s;
    return lo;
  };
  uint32_t sA = 0, lA = 0, sB = 0, lB = 0, sC = 0, lC = 0, sD = 0, lD = 0;
  // runs in descending key order (D, C, B, A), each search bounded by the previous run's position
  uint32_t hi = u;
  if (cy > 0) hi = range(hi, key - w0 - xl, key - w0 + xr, sD, lD);
  if (cz > 0) {
    const unsigned long long kb = key - w0 * w1;  // same (cx, cy), plane below
    if (cy + 1 < w1) hi = range(hi, kb + w0 - xl, kb + w0 + xr, sC, lC);
    hi = range(hi, kb - xl, kb + xr, sB, lB);
    if (cy > 0) hi = range(hi, kb - w0 - xl, kb - w0 + xr, sA, lA);
  }
  const uint32_t sE = (xl && u > 0 && __ldg(p.ukeys + u - 1) == key - 1) ? __ldg(p.csr + u - 1) : hb, lE = he - sE;
  r.o1 = lA; r.o2 = r.o1 + lB; r.o3 = r.o2 + lC; r.o4 = r.o3 + lD; r.K = r.o4 + lE;
  r.shA = sA; r.shB = sB - r.o1; r.shC = sC - r.o2; r.shD = sD - r.o3; r.shE = sE - r.o4;
  return true;
}

// ---------------------------------------------------------------------------------------------
// Exact loop: one warp enumerates the half-shell pairs of home cell c in the arithmetic of T.
//   recb / csrb are biased bases (GlobalRecs / StagedRecs): recb.load(pos) is record `pos` of the cell-sorted array and
//   csrb[cell] its CSR entry, whether they live in shared memory (staged tile) or in global memory.
//   Every lane keeps NJ candidates j in registers; the home particles i are broadcast loads.
// P > 1 is the packed tail: the chunk has at most 32 / P candidates, the warp is split into P
// phases (lane / (32 / P)) and phase ph tests home particles ph, ph + P, ...: the loop runs
// ceil(m / P) times instead of m.  A phase whose particle index runs past the cell reads a
// neighbouring record (in bounds: the stage and the record array have slack) and discards it
// through i < thr <= m.
template <class T, int CMP, int NJ, int P, class Recs, class Consumer>
__device__ __forceinline__ void exact_tests(const Recs home, uint32_t m, uint32_t ph,
                                            const T (&xj)[NJ], const T (&yj)[NJ], const T (&zj)[NJ],
                                            const uint32_t (&lj)[NJ], const uint32_t (&thr)[NJ], T c2,
                                            Consumer& cons) {
  constexpr int F = Consumer::kFuse;  // home particles per consumer call
  auto test = [&](uint32_t i, bool* h, T* dsq, uint32_t& li) {
    T xi, yi, zi;
    home.template load<Consumer::kNeedLabels>(i, xi, yi, zi, li);
#pragma unroll
    for (int q = 0; q < NJ; ++q) {
      h[q] = i < thr[q];
      dsq[q] = T(0);
      if (CMP != 0) {
        const T dx = xi - xj[q], dy = yi - yj[q], dz = zi - zj[q];
        dsq[q] = (dx * dx + dy * dy) + dz * dz;
        h[q] = h[q] && passes<CMP>(dsq[q], c2);
      }
    }
  };
  // F home particles per consumer call: their distance tests are independent instruction chains
  // that overlap, and the consumer's bookkeeping is paid once.  Indices past the cell (i >= m >=
  // thr) never hit; the reads stay in bounds (see above).
  uint32_t ljf[NJ * F];
#pragma unroll
  for (int f = 0; f < F; ++f)
#pragma unroll
    for (int q = 0; q < NJ; ++q) ljf[f * NJ + q] = lj[q];
  // f32 grids: two tests per packed instruction -- the two candidates of a lane (NJ = 2) or the two fused
  // home particles (NJ = 1, F = 2).  Same arithmetic, same roundings: (dx*dx + dy*dy) + dz*dz per element.
  constexpr bool kPacked = sizeof(T) == 4 && CMP != 0 && (NJ == 2 || (NJ == 1 && F == 2));
  if constexpr (kPacked) {
    const uint64_t one2 = pk2(1.0f, 1.0f), mone2 = pk2(-1.0f, -1.0f), zero2 = pk2(0.0f, 0.0f);
    uint64_t cx, cy, cz;  // NJ = 2: the lane's candidate pair
    if (NJ == 2) {
      cx = pk2((float)xj[0], (float)xj[NJ - 1]);
      cy = pk2((float)yj[0], (float)yj[NJ - 1]);
      cz = pk2((float)zj[0], (float)zj[NJ - 1]);
    }
    auto dsq2 = [&](uint64_t ax, uint64_t ay, uint64_t az, uint64_t bx, uint64_t by, uint64_t bz, float& d0, float& d1) {
      const uint64_t dx = ffma2(bx, mone2, ax), dy = ffma2(by, mone2, ay), dz = ffma2(bz, mone2, az);  // a - b
      const uint64_t sx = ffma2(dx, dx, zero2), sy = ffma2(dy, dy, zero2), sz = ffma2(dz, dz, zero2);
      const uint64_t s = ffma2(ffma2(sx, one2, sy), one2, sz);
      uint32_t u0, u1;
      unpk2(s, u0, u1);
      d0 = __uint_as_float(u0);
      d1 = __uint_as_float(u1);
    };
#pragma unroll Consumer::kUnroll
    for (uint32_t ib = 0; ib < m; ib += F * P) {
      bool h[NJ * F];
      T dsq[NJ * F];
      uint32_t lif[NJ * F];
      float xi[F], yi[F], zi[F];
      uint32_t idx[F];
#pragma unroll
      for (int f = 0; f < F; ++f) {
        uint32_t li;
        T x, y, z;
        idx[f] = (P == 1 ? ib : ib + ph) + f * P;
        home.template load<Consumer::kNeedLabels>(idx[f], x, y, z, li);
        xi[f] = (float)x; yi[f] = (float)y; zi[f] = (float)z;
#pragma unroll
        for (int q = 0; q < NJ; ++q) lif[f * NJ + q] = li;
      }
      if (NJ == 2) {
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float d0, d1;
          dsq2(pk2(xi[f], xi[f]), pk2(yi[f], yi[f]), pk2(zi[f], zi[f]), cx, cy, cz, d0, d1);
          dsq[f * NJ] = (T)d0;
          dsq[f * NJ + NJ - 1] = (T)d1;
        }
      } else {
        float d0, d1;
        dsq2(pk2(xi[0], xi[F - 1]), pk2(yi[0], yi[F - 1]), pk2(zi[0], zi[F - 1]), pk2((float)xj[0], (float)xj[0]),
             pk2((float)yj[0], (float)yj[0]), pk2((float)zj[0], (float)zj[0]), d0, d1);
        dsq[0] = (T)d0;
        dsq[NJ * F - 1] = (T)d1;
      }
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int q = 0; q < NJ; ++q) h[f * NJ + q] = idx[f] < thr[q] && passes<CMP>(dsq[f * NJ + q], c2);
      cons.template test_n<NJ * F>(h, dsq, lif, ljf);
    }
    return;
  }
#pragma unroll Consumer::kUnroll
  for (uint32_t ib = 0; ib < m; ib += F * P) {
    bool h[NJ * F];
    T dsq[NJ * F];
    uint32_t lif[NJ * F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      uint32_t li;
      test((P == 1 ? ib : ib + ph) + f * P, h + f * NJ, dsq + f * NJ, li);
#pragma unroll
      for (int q = 0; q < NJ; ++q) lif[f * NJ + q] = li;
    }
    cons.template test_n<NJ * F>(h, dsq, lif, ljf);
  }
}

// one group of up to 32 NJ candidates (lane l holds candidates kb + 32 q + l), or -- P > 1 -- a
// packed tail of up to 32 / P candidates (lane l holds candidate kb + l % (32 / P))
template <class T, int CMP, int NJ, int P, class Recs, class Consumer>
__device__ __forceinline__ void process_group(const CellRuns& r, const Recs recb, uint32_t kb, T c2,
                                              Consumer& cons) {
  constexpr uint32_t W = 32 / P;
  const unsigned lane = lane_id();
  const uint32_t slot = P == 1 ? lane : (lane & (W - 1)), ph = P == 1 ? 0u : lane / W;
  T xj[NJ], yj[NJ], zj[NJ];
  uint32_t lj[NJ], thr[NJ];
#pragma unroll
  for (int q = 0; q < NJ; ++q) {
    const uint32_t k = kb + 32u * q + slot;
    thr[q] = r.thr(k);
    recb.template load<Consumer::kNeedLabels>(r.pos(k), xj[q], yj[q], zj[q], lj[q]);
  }
  exact_tests<T, CMP, NJ, P>(recb.at(r.hb), r.m, ph, xj, yj, zj, lj, thr, c2, cons);
  cons.chunk_end();
}

template <class T, int CMP, class Recs, class Consumer>
__device__ __forceinline__ void process_cell(const CellRuns& r, const Recs recb, T c2, Consumer& cons) {
  constexpr int NJMAX = GenericNJ<T>::value;
  if (r.m == 0) return;
  if (CMP == 0 && Consumer::kCountsOnly) {  // unfiltered count: no per-pair work
    const unsigned lane = lane_id();
    uint32_t sum_thr = 0;
    for (uint32_t kb = 0; kb < r.K; kb += 32) sum_thr += r.thr(kb + lane);
    cons.add(sum_thr);
    return;
  }
  uint32_t kb = 0;
  while (kb < r.K) {
    const uint32_t rem = r.K - kb;
    if (NJMAX >= 2 && rem > 48) {          // two full-ish slots: one broadcast load per 2 tests
      process_group<T, CMP, 2, 1>(r, recb, kb, c2, cons);
      kb += 64;
    } else if (rem > kTailSplit) {         // one slot (a 33..48 remainder leaves a packed tail behind)
      process_group<T, CMP, 1, 1>(r, recb, kb, c2, cons);
      kb += 32;
    } else if (rem > 8) {                  // two phases of 16 lanes: <= 16 candidates, or the first 16 of 17..24
                                           // (then <= 8 are left for four phases: 0.75 m iterations instead of m)
      process_group<T, CMP, 1, 2>(r, recb, kb, c2, cons);
      kb += 16;
    } else {                               // <= 8 candidates: four phases of 8 lanes
      process_group<T, CMP, 1, 4>(r, recb, kb, c2, cons);
      kb += 8;
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <class T, int CMP, class Consumer, int MODE>
__global__ void __launch_bounds__(kPairThreads, Consumer::kMinBlocks) pair_kernel(PairParams<T> p, typename Consumer::Args args) {
  constexpr bool kSparse = MODE == 3;  // compact cells of a sparse grid: global-memory tiles, searched descriptors
  constexpr bool kStagedOnly = MODE == 1, kGlobalOnly = MODE == 2 || kSparse;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Rec<T>* s_rec = reinterpret_cast<Rec<T>*>(smem_raw);
  CellRuns* s_desc = reinterpret_cast<CellRuns*>(s_rec + (kGlobalOnly ? 0u : p.stage_recs));
  uint32_t* s_csr = reinterpret_cast<uint32_t*>(s_desc + kMaxTileCells);
  unsigned char* s_cons = reinterpret_cast<unsigned char*>(s_csr + (kGlobalOnly ? 4 : kStageCells + 4));
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ ConsumerSmem s_cs;
  __shared__ uint32_t s_next;  // next unclaimed home cell of the tile (dynamic balance between warps)
  __shared__ uint32_t s_tile[2];  // this CTA's next work item (claimed one tile ahead)

  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const T c2 = keep_in_reg(p.c2);
  Consumer cons(args, &s_cs, s_cons + (size_t)warp * Consumer::kWarpSmemBytes, c2);

  if (!kGlobalOnly && threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  uint32_t phase = 0;

  const uint32_t plane = (uint32_t)p.w0 * (uint32_t)p.w1;
  const uint32_t halo = plane + (uint32_t)p.w0 + 1u;
  const uint32_t nwork = p.work_list ? *p.work_list_n : (p.tile_list ? __ldg(p.tile_list_n) : p.ntiles);
  // Work items are claimed DYNAMICALLY: the first wave statically (blockIdx.x), every further one
  // from a global counter, so that CTAs which drew cheap tiles take more of them (clustered
  // particle distributions; also the last, partial wave of a uniform box).  The claim is issued
  // at the start of a tile and read at its end, its round trip hides behind the tile's work;
  // s_tile is double-buffered so that the next claim cannot overwrite a value still being read.
  uint32_t par = 0;
  for (uint32_t wi = blockIdx.x; wi < nwork;) {
    if (threadIdx.x == 0) s_tile[par] = gridDim.x + atomicAdd(p.tile_next, 1u);
    const uint32_t w = p.work_list ? p.work_list[wi] : wi;  // plain loads: written by the launch before this one
    const uint32_t tile = p.tile_list ? __ldg(p.tile_list + w) : w;
    uint32_t c0, c1;
    tile_cells_of(tile, p.home_lo, p.home_hi, p.tile_cells, kSparse ? 0u : p.row_tiles, (uint32_t)p.w0, c0, c1);
    const bool rows = !kSparse && !kGlobalOnly && p.row_tiles != 0u;  // five row segments instead of one range
    const uint32_t cl = (kSparse || rows) ? c0 : (c0 > halo ? c0 - halo : 0u);  // (sparse grids: no staged halo, `halo` may have wrapped)
    const uint32_t ncsr = c1 - cl + 1;
    uint32_t plo = __ldg(p.csr + cl);
    const uint32_t phi = __ldg(p.csr + c1);
    uint32_t np = phi - plo;
    // row tiles: the five ranges (A, B, C in the plane below, D the row before, E the home row with its left
    // neighbour cell), as record ranges [rs[x], re[x]) and their offsets ro[x] in the stage
    uint32_t rs[5] = {0, 0, 0, 0, 0}, re[5] = {0, 0, 0, 0, 0}, ro[5] = {0, 0, 0, 0, 0};
    if (rows) {
      const uint32_t w0 = (uint32_t)p.w0, w1 = (uint32_t)p.w1;
      const uint32_t row = fast_div(c0, p.div0), cx0 = c0 - row * w0, cx1 = cx0 + (c1 - c0);  // x range [cx0, cx1)
      const uint32_t cz = fast_div(row, p.div1), cy = row - cz * w1;
      const uint32_t xa = cx0 > 0 ? cx0 - 1 : 0u, xb = min(cx1 + 1, w0);  // with the x neighbours
      const uint32_t base = row * w0;
      auto seg = [&](int x, bool on, uint32_t rbase, uint32_t hi_x) {
        if (on) {
          rs[x] = __ldg(p.csr + rbase + xa);
          re[x] = __ldg(p.csr + rbase + hi_x);
        }
      };
      seg(0, cz > 0 && cy > 0, base - plane - w0, xb);
      seg(1, cz > 0, base - plane, xb);
      seg(2, cz > 0 && cy + 1 < w1, base - plane + w0, xb);
      seg(3, cy > 0, base - w0, xb);
      seg(4, true, base, cx1);
      np = 0;
#pragma unroll
      for (int x = 0; x < 5; ++x) {
        ro[x] = np;
        np += re[x] - rs[x];
      }
      plo = 0;  // records are addressed through the per-run offsets below
    }
    const bool staged = !kGlobalOnly && np <= p.stage_recs && (rows || ncsr <= (uint32_t)kStageCells);
    // sparse boxes: a tile without home particles has no pairs (its per-tile count stays 0).
    // Staged-only kernel: a tile that does not fit the stage goes to the list of the global-memory launch.
    const bool empty = phi == __ldg(p.csr + c0);
    if (empty || (kStagedOnly && !staged)) {
      if (kStagedOnly && !empty && threadIdx.x == 0) p.fb_list[atomicAdd(p.fb_count, 1u)] = w;
      __syncthreads();
      wi = s_tile[par];
      par ^= 1u;
      continue;
    }

    cons.tile_begin(w, s_rec, false);  // per-tile arrays are indexed by work item (= tile id unless sparse)
    if (threadIdx.x == 0) s_next = c0 + kPairWarps;
    if (staged) {
      if (threadIdx.x == 0 && np > 0) {
        mbar_expect_tx(&s_bar, np * (uint32_t)sizeof(Rec<T>));
        if (rows) {
#pragma unroll
          for (int x = 0; x < 5; ++x)
            if (re[x] > rs[x]) bulk_g2s(s_rec + ro[x], p.sorted + rs[x], (re[x] - rs[x]) * (uint32_t)sizeof(Rec<T>), &s_bar);
        } else {
          bulk_g2s(s_rec, p.sorted + plo, np * (uint32_t)sizeof(Rec<T>), &s_bar);
        }
      }
      if (!rows) {
        for (uint32_t k = threadIdx.x; k < ncsr; k += kPairThreads) s_csr[k] = __ldg(p.csr + cl + k);
        __syncthreads();
      }
    }
    // one thread per home cell computes the cell's run descriptor while the bulk copy is in flight
    {
      const uint32_t* csrb = (staged && !rows) ? s_csr - cl : p.csr;
      for (uint32_t k = threadIdx.x; k < c1 - c0; k += kPairThreads) {
        CellRuns r;
        if constexpr (kSparse) cell_runs_sparse(p, c0 + k, r);
        else cell_runs(p, c0 + k, csrb, r);
        if (rows && staged) {
          // record index -> stage index: every run lives in its own segment of the stage
          r.shA += ro[0] - rs[0];
          r.shB += ro[1] - rs[1];
          r.shC += ro[2] - rs[2];
          r.shD += ro[3] - rs[3];
          r.shE += ro[4] - rs[4];
          r.hb += ro[4] - rs[4];
        }
        s_desc[k] = r;
      }
    }
    if (staged && np > 0) {
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
    }
    __syncthreads();
    // the stage as a biased shared-space address, pinned in a register for the whole tile
    StagedRecs<T> srecs{smem_u32(s_rec) - plo * (uint32_t)sizeof(Rec<T>)};
    asm volatile("" : "+r"(srecs.a));
    // warps claim home cells one at a time: the first kPairWarps statically, the rest from s_next
    for (uint32_t c = c0 + warp; c < c1;) {
      const CellRuns r = s_desc[c - c0];
      // separate call sites so that the staged ones compile to shared-memory loads (LDS)
      if constexpr (kStagedOnly) {
        process_cell<T, CMP>(r, srecs, c2, cons);
      } else if constexpr (kGlobalOnly) {
        process_cell<T, CMP>(r, GlobalRecs<T>{p.sorted}, c2, cons);
      } else {
        if (staged) process_cell<T, CMP>(r, srecs, c2, cons);
        else process_cell<T, CMP>(r, GlobalRecs<T>{p.sorted}, c2, cons);
      }
      uint32_t nxt = 0;
      if (lane == 0) nxt = atomicAdd(&s_next, 1u);
      c = __shfl_sync(0xffffffffu, nxt, 0);
    }
    cons.template tile_end<CMP>(w);  // ends with __syncthreads(): the stage buffers may be overwritten
    wi = s_tile[par];
    par ^= 1u;
  }
  cons.finish();
  // every claim of this CTA precedes its `done` tick, so the CTA that sees the last tick knows that
  // nobody will touch tile_next again (atomicInc wraps tile_done back to 0 by itself)
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(p.tile_done, gridDim.x - 1) == gridDim.x - 1) {
      *p.tile_next = 0u;
      if (p.work_list) *p.work_list_n = 0u;  // every CTA has read its bound: empty list for the next launch
    }
  }
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

// sparse boxes: the tiles whose home cells hold at least one particle (order unspecified)
__global__ void tile_list_kernel(const uint32_t* __restrict__ csr, uint32_t home_lo, uint32_t home_hi,
                                 uint32_t tile_cells, uint32_t row_tiles, uint32_t w0, uint32_t ntiles,
                                 uint32_t* __restrict__ list, uint32_t* __restrict__ count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  bool live = false;
  if (t < ntiles) {
    uint32_t c0, c1;
    tile_cells_of(t, home_lo, home_hi, tile_cells, row_tiles, w0, c0, c1);
    live = __ldg(csr + c1) != __ldg(csr + c0);
  }
  const unsigned b = __ballot_sync(0xffffffffu, live);
  if (b == 0) return;
  uint32_t base = 0;
  if (lane_id() == 0) base = atomicAdd(count, (uint32_t)__popc(b));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (live) list[base + __popc(b & lanemask_lt())] = t;
}

// (energy, pair count, error flag = 0) -> three doubles for one all-reduce (the count is exact in f64
// below 2^53; a rank whose step failed contributes (0, 0, 1) from the host instead)
// flags / slab_flag: the build's own verdict (particle outside the window; halo overflow; box changed under a
// speculative step) is folded in on the device, so no host round trip is needed before the collective.
__global__ void pack_energy_count_kernel(const double* __restrict__ e, const unsigned long long* __restrict__ c,
                                         const int* __restrict__ flags, const uint32_t* __restrict__ slab_flag,
                                         double* __restrict__ out3) {
  out3[0] = *e;
  out3[1] = (double)*c;
  out3[2] = ((*flags & 1) || (*slab_flag & 15u)) ? 1.0 : 0.0;
}

// exclusive scan of the per-tile pair counts (a few 10^4 entries): one block, serial over chunks
__global__ void tile_offsets_kernel(const unsigned long long* __restrict__ counts, uint32_t ntiles,
                                    const uint32_t* __restrict__ ntiles_dev, unsigned long long* __restrict__ offsets) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  if (ntiles_dev) ntiles = *ntiles_dev;  // sparse boxes: number of listed tiles
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
