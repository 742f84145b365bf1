// pair_pf_kernels.cuh -- the f64 pair pass with an f32 prefilter (default for the pair COUNT of f64 grids;
// ZB_PREFILTER selects the consumers, DESIGN.md section 5b has the measurements behind that default).
//
// Same enumeration as pair_kernels.cuh (GridCell::particle_pairs, iters.rs:238-241: z-major half shell =
// 5 runs of consecutive records per home cell, one warp per home cell, lanes = candidates), same
// consumers, same bit-exact results -- but the 8.2e8 distance tests of the benchmark no longer run on
// the FP64 pipe (16 lanes / clk / SM sub-partition: profiles/fp64_peak.json puts the reference's exact
// 9-operation test at 1.6e12 tests/s for the whole GPU, a floor of 0.6 ms at n = 10^7) and hits are no
// longer compacted test by test (ballot + 2 popc + 2 mad + predicated store per test in the exact kernel).
//
//  1. A tile's records arrive by ONE TMA bulk copy (cp.async.bulk + mbarrier) into a landing buffer and
//     are re-laid by all threads into structure-of-arrays form: exact f64 coordinates (xd, yd, zd) and
//     f32 coordinates RELATIVE to the tile's first record (xf, yf, zf).  The landing buffer then serves
//     as the warps' hit queues.
//  2. Test loop, packed f32x2 (fma.rn.f32x2, SASS FFMA2): a lane holds up to 4 candidates j; the home
//     particles come TWO per step as natural pairs (x_i, x_i+1) = one aligned 8-byte broadcast load from
//     the f32 array.  Per 2 tests: 3 FFMA2 for the differences, 3 FFMA2 for
//     t = dx^2 + dy^2 + dz^2 - hi, and the SIGN of t is shifted into a per-candidate bit mask by one
//     funnel shift per test (no compare, no ballot, no predicate in the loop).  Which home particles a
//     candidate may pair with (all of them; only the earlier ones for a candidate of the home cell
//     itself, iters.rs:29-36; none for an idle lane) is applied to the mask afterwards.
//  3. `hi` = c^2 (1 + delta) with the guard band of prefilter_delta(): t >= 0 proves the pair is outside
//     the cutoff.  The count consumer adds up the set bits; pairs whose t lies below -(hi - lo) are
//     certainly inside, and it re-decides a lane's pairs in f64 only when the signed-integer minimum over
//     the lane's t bit patterns (one 3-input VIMNMX per two tests) shows a value inside the band.
//  4. The other consumers decide EVERY set bit ("maybe") in f64 with the reference's own arithmetic
//     ((dx*dx + dy*dy) + dz*dz, separately rounded, `<` / `<=` against cutoff.powi(2)): once per cell a
//     warp prefix sum over the lanes' bit counts gives every lane its place in the warp's queue, the
//     lanes push (home record, candidate record) entries, and full rows of 32 entries are evaluated with
//     all lanes busy.  Entries are in lane order, so a row's candidates are (mostly) consecutive records
//     and its home particles the ~10 records of one cell: the f64 gathers are nearly conflict-free.
//
// Tiles the stage cannot hold, tiles whose f32 guard band would be too wide (delta >= 0.25) and tiles
// with non-finite coordinates are appended to a work list; the exact kernel of pair_kernels.cuh runs
// over that list afterwards (launch_pairs in zelll_b200.cu).
#pragma once

#include "pair_kernels.cuh"

namespace zb {

#ifndef ZB_PF_STAGE_RECS
#define ZB_PF_STAGE_RECS 896
#endif
#ifndef ZB_PF_MINBLOCKS
#define ZB_PF_MINBLOCKS 3
#endif
constexpr uint32_t kPfStageRecs = ZB_PF_STAGE_RECS;  // records per stage (multiple of 8); 3 CTAs per SM at 896
constexpr uint32_t kPfSP = kPfStageRecs + 8;        // array pitch: slack for the pair loads around a cell
constexpr int kPfMaxNJ = 4;                         // candidates per lane (register tile)
constexpr float kPfIdle = 1.0e18f;    // coordinate of an idle lane's "candidate": t = +huge, never a hit
constexpr float kPfMaxRel = 1.0e15f;  // |relative coordinate| beyond this (or NaN) sends the tile to the exact kernel
constexpr uint32_t kPfQueueCap = kPfStageRecs * (uint32_t)sizeof(Rec<double>) / 4u / kPairWarps;  // entries per warp
static_assert(kPfStageRecs % 8 == 0 && kPfQueueCap >= 256, "stage size");

// shared-memory loads at (address register + compile-time byte offset)
template <uint32_t OFF>
__device__ __forceinline__ uint64_t lds_b64(uint32_t a) {
  uint64_t v;
  asm volatile("ld.shared.b64 %0, [%1+%2];" : "=l"(v) : "r"(a), "n"(OFF));
  return v;
}
template <uint32_t OFF>
__device__ __forceinline__ double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
  return v;
}
template <uint32_t OFF>
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// x << s with s clamped to 32 (PTX shl semantics; C++ leaves s >= 32 undefined)
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t s) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
  return r;
}

// shared-memory views of one staged tile (stage-local record index = record - plo).  The y / z arrays
// sit at fixed byte offsets behind the x arrays: one address register per access.
struct PfStage {
  uint32_t xf;  // shared-space byte address of the f32 x array (y at + 4 SP, z at + 8 SP)
  uint32_t xd;  // ... of the f64 x array (y at + 8 SP, z at + 16 SP)
  uint32_t lab; // ... of the label array (emit consumer)
  uint32_t q;   // ... of this warp's hit queue (kPfQueueCap entries of (home | candidate << 16))
  uint32_t plo;
};
constexpr uint32_t kPfOffF = kPfSP * 4u;  // byte pitch between the f32 arrays
constexpr uint32_t kPfOffD = kPfSP * 8u;  // ... between the f64 arrays

struct PfThresh {
  uint64_t nhi2;   // (-hi, -hi)
  int32_t band;    // bit pattern of -(hi - lo) as a signed integer: t bits <= band  <=>  t in [-(hi-lo), -0]
};

// ---------------------------------------------------------------------------------------------
// test loop: S steps of two home particles each, starting at stage-local (even) index b0
template <int NJ, bool TRACK_MIN>
__device__ __forceinline__ void pf_tests(const PfStage& st, uint32_t b0, uint32_t S, const uint64_t (&cx)[kPfMaxNJ],
                                         const uint64_t (&cy)[kPfMaxNJ], const uint64_t (&cz)[kPfMaxNJ],
                                         const PfThresh& th, uint32_t (&mask)[kPfMaxNJ], int32_t& tmin) {
  const uint64_t mone2 = pk2(-1.0f, -1.0f);
  uint32_t a = st.xf + b0 * 4u;
#pragma unroll 2
  for (uint32_t s = 0; s < S; ++s) {
    const uint64_t hx = lds_b64<0>(a), hy = lds_b64<kPfOffF>(a), hz = lds_b64<2 * kPfOffF>(a);
    a += 8u;
#pragma unroll
    for (int q = 0; q < NJ; ++q) {
      const uint64_t dx = ffma2(hx, mone2, cx[q]);  // x_j - x_i for both home particles of the step
      const uint64_t dy = ffma2(hy, mone2, cy[q]);
      const uint64_t dz = ffma2(hz, mone2, cz[q]);
      const uint64_t t = ffma2(dz, dz, ffma2(dy, dy, ffma2(dx, dx, th.nhi2)));
      uint32_t ta, tb;
      unpk2(t, ta, tb);
      mask[q] = __funnelshift_l(ta, mask[q], 1);  // sign of t = "maybe inside"
      mask[q] = __funnelshift_l(tb, mask[q], 1);
      if (TRACK_MIN) tmin = __vimin3_s32(tmin, (int32_t)ta, (int32_t)tb);
    }
  }
}

// the reference's exact test between stage-local records i and j
template <int CMP>
__device__ __forceinline__ bool pf_exact(const PfStage& st, uint32_t i, uint32_t j, double c2, double& dsq) {
  const uint32_t ai = st.xd + i * 8u, aj = st.xd + j * 8u;
  const double xi = lds_f64<0>(ai), yi = lds_f64<kPfOffD>(ai), zi = lds_f64<2 * kPfOffD>(ai);
  const double xj = lds_f64<0>(aj), yj = lds_f64<kPfOffD>(aj), zj = lds_f64<2 * kPfOffD>(aj);
  const double dx = xi - xj, dy = yi - yj, dz = zi - zj;
  dsq = (dx * dx + dy * dy) + dz * dz;
  return passes<CMP>(dsq, c2);
}

// one row of queue entries (home | candidate << 16), decided in f64 with all lanes busy
template <int CMP, class Consumer>
__device__ __forceinline__ void pf_row(const PfStage& st, uint32_t entry, bool valid, double c2, Consumer& cons) {
  const uint32_t i = entry & 0xffffu, j = entry >> 16;
  double dsq;
  const bool h = pf_exact<CMP>(st, i, j, c2, dsq) && valid;
  uint32_t li = 0, lj = 0;
  if (Consumer::kNeedLabels) {
    li = lds_u32(st.lab + i * 4u);
    lj = lds_u32(st.lab + j * 4u);
  }
  cons.hit(h, dsq, li, lj);
}

// ---------------------------------------------------------------------------------------------
// Count consumer: one chunk of up to 32 NJ candidates, everything specialised by NJ (a cell's last chunk
// rarely fills four slots; set-up and mask bookkeeping are paid per slot in use).
template <int CMP, int NJ, class Consumer>
__device__ __forceinline__ void pf_count_chunk(const CellRuns& r, const PfStage& st, uint32_t kb, const PfThresh& th, double c2,
                                               Consumer& cons) {
  const unsigned lane = lane_id();
  const uint32_t hbl = r.hb - st.plo, hend = hbl + r.m;
  const uint32_t first_home = r.K - r.m;  // candidates [first_home, K) are the home cell's own particles
  uint64_t cx[kPfMaxNJ], cy[kPfMaxNJ], cz[kPfMaxNJ];
  uint32_t posl[NJ], lim[NJ];
#pragma unroll
  for (int q = 0; q < NJ; ++q) {
    const uint32_t k = kb + 32u * q + lane;
    const bool valid = k < r.K;
    posl[q] = r.pos(k) - st.plo;
    lim[q] = valid ? (k >= first_home ? posl[q] : hend) : hbl;
    float x = kPfIdle, y = kPfIdle, z = kPfIdle;
    if (valid) {
      const uint32_t a = st.xf + posl[q] * 4u;
      x = lds_f32<0>(a);
      y = lds_f32<kPfOffF>(a);
      z = lds_f32<2 * kPfOffF>(a);
    }
    cx[q] = pk2(x, x);
    cy[q] = pk2(y, y);
    cz[q] = pk2(z, z);
  }
#pragma unroll 1
  for (uint32_t b0 = hbl & ~1u; b0 < hend; b0 += 32u) {
    const uint32_t S = min(16u, (hend - b0 + 1u) >> 1);
    uint32_t mask[kPfMaxNJ];
#pragma unroll
    for (int q = 0; q < kPfMaxNJ; ++q) mask[q] = 0u;
    int32_t tmin = 0x7fffffff;
    pf_tests<NJ, true>(st, b0, S, cx, cy, cz, th, mask, tmin);
    const uint32_t top = b0 + 2u * S - 1u;
    const uint32_t first = max(hbl, b0);
    const uint32_t himask = top - first >= 31u ? 0xffffffffu : (2u << (top - first)) - 1u;
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < NJ; ++q) {
      const int32_t jlo = (int32_t)(top + 1u) - (int32_t)lim[q];
      c += __popc(mask[q] & himask & shl_clamp(0xffffffffu, (uint32_t)max(jlo, 0)));
    }
    // some t of this lane inside the guard band (taken over ALL its tests, masked-out ones included:
    // conservative): decide the lane's pairs of this pass in f64
    const bool amb = tmin <= th.band;
    if (__any_sync(0xffffffffu, amb)) {
      if (amb) {
        c = 0;
        uint32_t pq[NJ], lq[NJ];
#pragma unroll
        for (int q = 0; q < NJ; ++q) { pq[q] = posl[q]; lq[q] = lim[q]; }
#pragma unroll 1
        for (int q = 0; q < NJ; ++q) {
          const uint32_t e = min(lq[0], top + 1u);
#pragma unroll 1
          for (uint32_t i = first; i < e; ++i) {
            double dsq;
            c += pf_exact<CMP>(st, i, pq[0], c2, dsq) ? 1u : 0u;
          }
#pragma unroll
          for (int u = 0; u + 1 < NJ; ++u) { pq[u] = pq[u + 1]; lq[u] = lq[u + 1]; }
        }
      }
    }
    cons.add(c);
  }
}

// ---------------------------------------------------------------------------------------------
// One home cell.  Its K candidates are taken in chunks of up to 128 (lane l holds candidates kb + 32 q + l,
// q < nj <= 4); only the test loop is specialised by nj, everything around it exists once per kernel
// (the kernel has to stay small: instruction cache).  `qn` = entries waiting in this warp's queue.
template <int CMP, class Consumer>
__device__ __forceinline__ void pf_cell(const CellRuns& r, const PfStage& st, const PfThresh& th, double c2,
                                        Consumer& cons, uint32_t& qn) {
  constexpr bool kCount = Consumer::kCountsOnly;
  constexpr int NJ = kPfMaxNJ;
  if (r.m == 0) return;
  if constexpr (kCount) {
#pragma unroll 1
    for (uint32_t kb = 0; kb < r.K; kb += 32u * NJ) {
      const uint32_t nj = (r.K - kb + 31u) >> 5;
      if (nj >= 4) pf_count_chunk<CMP, 4>(r, st, kb, th, c2, cons);
      else if (nj == 3) pf_count_chunk<CMP, 3>(r, st, kb, th, c2, cons);
      else if (nj == 2) pf_count_chunk<CMP, 2>(r, st, kb, th, c2, cons);
      else pf_count_chunk<CMP, 1>(r, st, kb, th, c2, cons);
    }
    return;
  }
  const unsigned lane = lane_id();
  const uint32_t hbl = r.hb - st.plo;  // stage-local index of the home cell's first record
  const uint32_t hend = hbl + r.m;
#pragma unroll 1
  for (uint32_t kb = 0; kb < r.K; kb += 32u * NJ) {
    const uint32_t nj = min((r.K - kb + 31u) >> 5, (uint32_t)NJ);
    uint64_t cx[NJ], cy[NJ], cz[NJ];
    uint32_t posl[NJ], lim[NJ];
#pragma unroll
    for (int q = 0; q < NJ; ++q) {
      posl[q] = hbl;
      lim[q] = hbl;  // idle slot: pairs with nothing
      float x = kPfIdle, y = kPfIdle, z = kPfIdle;
      if (q < (int)nj) {  // warp-uniform
        const uint32_t k = kb + 32u * q + lane;
        const bool valid = k < r.K;
        posl[q] = r.pos(k) - st.plo;  // idle lane: the home cell's first record (in bounds)
        // home particles [hbl, lim) pair with this candidate: all of them, or -- for a candidate that is
        // the home cell's own particle -- the ones stored before it (intra-cell pairs once, iters.rs:29-36)
        lim[q] = valid ? (k >= r.K - r.m ? posl[q] : hend) : hbl;
        if (valid) {
          const uint32_t a = st.xf + posl[q] * 4u;
          x = lds_f32<0>(a);
          y = lds_f32<kPfOffF>(a);
          z = lds_f32<2 * kPfOffF>(a);
        }
      }
      cx[q] = pk2(x, x);
      cy[q] = pk2(y, y);
      cz[q] = pk2(z, z);
    }
    // home particles in passes of up to kPass (aligned pairs; 16 so that two masks share a word, 32 for
    // the count consumer); the first pass starts one record early when the cell starts at an odd index
    // (that record's bit is masked out below)
    constexpr uint32_t kPass = kCount ? 32u : 16u;
#pragma unroll 1
    for (uint32_t b0 = hbl & ~1u; b0 < hend; b0 += kPass) {
      const uint32_t S = min(kPass / 2u, (hend - b0 + 1u) >> 1);
      uint32_t mask[NJ];
#pragma unroll
      for (int q = 0; q < NJ; ++q) mask[q] = 0u;
      int32_t tmin = 0x7fffffff;
      if (nj == 1) pf_tests<1, kCount>(st, b0, S, cx, cy, cz, th, mask, tmin);
      else if (nj == 2) pf_tests<2, kCount>(st, b0, S, cx, cy, cz, th, mask, tmin);
      else if (nj == 3) pf_tests<3, kCount>(st, b0, S, cx, cy, cz, th, mask, tmin);
      else pf_tests<4, kCount>(st, b0, S, cx, cy, cz, th, mask, tmin);
      // bit j of a mask belongs to home record top - j; keep records in [max(hbl, b0), lim)
      const uint32_t top = b0 + 2u * S - 1u;
      const uint32_t first = max(hbl, b0);
      const uint32_t himask = top - first >= 31u ? 0xffffffffu : (2u << (top - first)) - 1u;
#pragma unroll
      for (int q = 0; q < NJ; ++q) {
        const int32_t jlo = (int32_t)(top + 1u) - (int32_t)lim[q];
        mask[q] &= himask & shl_clamp(0xffffffffu, (uint32_t)max(jlo, 0));
      }
      if constexpr (kCount) {
        uint32_t c = (__popc(mask[0]) + __popc(mask[1])) + (__popc(mask[2]) + __popc(mask[3]));
        // some t of this lane inside the guard band (taken over ALL its tests, masked-out ones included:
        // conservative): decide the lane's pairs of this pass in f64
        const bool amb = tmin <= th.band;
        if (__any_sync(0xffffffffu, amb)) {
          if (amb) {
            c = 0;
            uint32_t pq[NJ], lq[NJ];
#pragma unroll
            for (int q = 0; q < NJ; ++q) { pq[q] = posl[q]; lq[q] = lim[q]; }
#pragma unroll 1
            for (int q = 0; q < NJ; ++q) {
              const uint32_t e = min(lq[0], top + 1u);
#pragma unroll 1
              for (uint32_t i = first; i < e; ++i) {
                double dsq;
                c += pf_exact<CMP>(st, i, pq[0], c2, dsq) ? 1u : 0u;
              }
#pragma unroll
              for (int u = 0; u + 1 < NJ; ++u) { pq[u] = pq[u + 1]; lq[u] = lq[u + 1]; }
            }
          }
        }
        cons.add(c);
      } else {
        // every "maybe" is decided in f64.  Entries go to the warp's queue in LANE order (prefix sum of
        // the lanes' bit counts); full rows of 32 are evaluated with all lanes busy.
        // two 32-bit words: bits [16 q', 16 q' + 16) of word w belong to candidate slot 2 w + q'
        uint32_t w0 = mask[0] | (mask[1] << 16), w1 = mask[2] | (mask[3] << 16);
        const uint32_t c = __popc(w0) + __popc(w1);
        const uint32_t total = __reduce_add_sync(0xffffffffu, c);
        if (total == 0) continue;
        const uint32_t e0 = top | (posl[0] << 16), e1 = top | (posl[1] << 16), e2 = top | (posl[2] << 16),
                       e3 = top | (posl[3] << 16);
        if (qn + total <= kPfQueueCap) {
          uint32_t incl = c;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += v;
          }
          uint32_t a = st.q + (qn + incl - c) * 4u;
          qn += total;
          // push: one entry per set bit, word 0 first; entry = (top - j) | candidate << 16
          const uint32_t its = __reduce_max_sync(0xffffffffu, c);
#pragma unroll 1
          for (uint32_t it = 0; it < its; ++it) {
            const bool first_word = w0 != 0u;
            const uint32_t wsel = first_word ? w0 : w1;
            if (wsel != 0u) {
              const uint32_t b = 31u - (uint32_t)__clz(wsel);
              const uint32_t rest = wsel ^ (1u << b);
              const uint32_t lo_e = first_word ? e0 : e2, hi_e = first_word ? e1 : e3;
              sts_u32(a, ((b & 16u) ? hi_e : lo_e) - (b & 15u));
              a += 4u;
              if (first_word) w0 = rest;
              else w1 = rest;
            }
          }
          __syncwarp();
#pragma unroll 1
          while (qn >= 32u) {
            qn -= 32u;
            pf_row<CMP>(st, lds_u32(st.q + (qn + lane) * 4u), true, c2, cons);
          }
          __syncwarp();
        } else {
          // more maybes than the queue holds (very dense cells): each lane decides its own, uncompacted
          const uint32_t its = __reduce_max_sync(0xffffffffu, c);
#pragma unroll 1
          for (uint32_t it = 0; it < its; ++it) {
            uint32_t entry = 0;
            const bool act = (w0 | w1) != 0u;
            if (w0 != 0u) {
              const uint32_t b = 31u - (uint32_t)__clz(w0);
              w0 ^= 1u << b;
              entry = ((b & 16u) ? e1 : e0) - (b & 15u);
            } else if (w1 != 0u) {
              const uint32_t b = 31u - (uint32_t)__clz(w1);
              w1 ^= 1u << b;
              entry = ((b & 16u) ? e3 : e2) - (b & 15u);
            }
            pf_row<CMP>(st, entry, act, c2, cons);
          }
        }
      }
    }
  }
}

// the entries still queued at the end of a tile (they name stage-local records)
template <int CMP, class Consumer>
__device__ __forceinline__ void pf_flush(const PfStage& st, double c2, Consumer& cons, uint32_t& qn) {
  if constexpr (!Consumer::kCountsOnly) {
    __syncwarp();
    if (qn > 0u) {  // < 32 entries
      const bool v = lane_id() < qn;
      pf_row<CMP>(st, v ? lds_u32(st.q + lane_id() * 4u) : 0u, v, c2, cons);
      qn = 0u;
    }
  }
}

// shared memory of one CTA: landing buffer (later the hit queues), 3 f64 + 3 f32 arrays, labels,
// descriptors, CSR slice, the consumers' per-warp bytes
__host__ __device__ constexpr size_t pf_smem_bytes(size_t warp_smem) {
  return (size_t)kPfStageRecs * sizeof(Rec<double>) + (size_t)kPfSP * (3 * 8 + 3 * 4 + 4) + kMaxTileCells * sizeof(CellRuns) +
         (kStageCells + 4) * sizeof(uint32_t) + kPairWarps * warp_smem;
}

// ---------------------------------------------------------------------------------------------
template <int CMP, class Consumer>
__global__ void __launch_bounds__(kPairThreads, ZB_PF_MINBLOCKS) pf_pair_kernel(PairParams<double> p,
                                                                               typename Consumer::Args args) {
  static_assert(CMP == 1 || CMP == 2, "the prefilter needs a distance filter");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr uint32_t SR = kPfStageRecs, SP = kPfSP;
  Rec<double>* s_land = reinterpret_cast<Rec<double>*>(smem_raw);
  double* s_xd = reinterpret_cast<double*>(s_land + SR);  // y at + SP, z at + 2 SP
  float* s_xf = reinterpret_cast<float*>(s_xd + 3 * SP);  // likewise
  uint32_t* s_lab = reinterpret_cast<uint32_t*>(s_xf + 3 * SP);
  CellRuns* s_desc = reinterpret_cast<CellRuns*>(s_lab + SP);
  uint32_t* s_csr = reinterpret_cast<uint32_t*>(s_desc + kMaxTileCells);
  unsigned char* s_cons = reinterpret_cast<unsigned char*>(s_csr + kStageCells + 4);
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ ConsumerSmem s_cs;
  __shared__ uint32_t s_next;
  __shared__ uint32_t s_tile[2];

  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const double c2 = keep_in_reg(p.c2);
  Consumer cons(args, &s_cs, s_cons + (size_t)warp * Consumer::kPfWarpSmemBytes, c2);

  if (threadIdx.x == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  uint32_t phase = 0;

  PfStage st;
  st.xf = smem_u32(s_xf);
  st.xd = smem_u32(s_xd);
  st.lab = smem_u32(s_lab);
  st.q = smem_u32(s_land) + (uint32_t)warp * kPfQueueCap * 4u;
  st.plo = 0;
  uint32_t qn = 0;

  const uint32_t plane = (uint32_t)p.w0 * (uint32_t)p.w1;
  const uint32_t halo = plane + (uint32_t)p.w0 + 1u;
  const uint32_t nwork = p.tile_list ? __ldg(p.tile_list_n) : p.ntiles;
  uint32_t par = 0;
  for (uint32_t w = blockIdx.x; w < nwork;) {
    if (threadIdx.x == 0) s_tile[par] = gridDim.x + atomicAdd(p.tile_next, 1u);
    const uint32_t tile = p.tile_list ? __ldg(p.tile_list + w) : w;
    const uint32_t c0 = p.home_lo + tile * p.tile_cells;
    const uint32_t c1 = min(c0 + p.tile_cells, p.home_hi);
    const uint32_t cl = c0 > halo ? c0 - halo : 0u;
    const uint32_t ncsr = c1 - cl + 1;
    const uint32_t plo = __ldg(p.csr + cl), phi = __ldg(p.csr + c1);
    const uint32_t np = phi - plo;
    const bool skip = phi == __ldg(p.csr + c0);  // no home particles: no pairs
    // guard band of this tile (prefilter_delta): R bounds |x - origin| over the staged cells
    float hi = 0.f, lo = 0.f;
    bool exact_only = np > SR || ncsr > (uint32_t)kStageCells;
    if (!skip && !exact_only) {
      const uint32_t nt = c1 - cl;
      const uint32_t sx = min((uint32_t)p.w0, nt);
      const uint32_t sy = min((uint32_t)p.w1, (nt + (uint32_t)p.w0 - 1) / (uint32_t)p.w0 + 1);
      const uint32_t sz = min((uint32_t)p.w2, (nt + plane - 1) / plane + 1);
      const float R = (float)max(sx, max(sy, sz)) * (float)p.cell * 1.0001f;
      const float delta = prefilter_delta(R / (float)p.fc);
      if (delta < 0.25f) {
        lo = __fmul_rd(__double2float_rd(c2), 1.0f - delta);
        hi = __fmul_ru(__double2float_ru(c2), 1.0f + delta);
      } else {
        exact_only = true;
      }
    }
    bool bad = false;
    if (!skip && !exact_only) {
      cons.tile_begin(w, s_land, false);
      if (threadIdx.x == 0) {
        s_next = c0 + kPairWarps;
        const uint32_t bytes = np * (uint32_t)sizeof(Rec<double>);
        // the landing buffer held the previous tile's queues (generic-proxy stores): order them before
        // the async-proxy write of the bulk copy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&s_bar, bytes);
        bulk_g2s(s_land, p.sorted + plo, bytes, &s_bar);
      }
      for (uint32_t k = threadIdx.x; k < ncsr; k += kPairThreads) s_csr[k] = __ldg(p.csr + cl + k);
      __syncthreads();
      {
        const uint32_t* csrb = s_csr - cl;
        for (uint32_t k = threadIdx.x; k < c1 - c0; k += kPairThreads) {
          CellRuns r;
          cell_runs(p, c0 + k, csrb, r);
          s_desc[k] = r;
        }
      }
      mbar_wait(&s_bar, phase);
      phase ^= 1u;
      // re-lay the records: exact f64 arrays + f32 coordinates relative to the stage's first record
      const double ox = s_land[0].x, oy = s_land[0].y, oz = s_land[0].z;
      for (uint32_t k = threadIdx.x; k < np + 8u; k += kPairThreads) {
        double x = ox, y = oy, z = oz;  // slack records: the origin itself
        uint32_t l = 0;
        if (k < np) load_part<Consumer::kNeedLabels>(s_land + k, x, y, z, l);
        s_xd[k] = x; s_xd[SP + k] = y; s_xd[2 * SP + k] = z;
        const float fx = (float)(x - ox), fy = (float)(y - oy), fz = (float)(z - oz);
        s_xf[k] = fx; s_xf[SP + k] = fy; s_xf[2 * SP + k] = fz;
        if (Consumer::kNeedLabels) s_lab[k] = l;
        bad = bad || !(fabsf(fx) < kPfMaxRel) || !(fabsf(fy) < kPfMaxRel) || !(fabsf(fz) < kPfMaxRel);
      }
      bad = __syncthreads_or(bad) != 0;  // also: the landing buffer is free to hold the queues from here on
    }
    if (skip) {
      __syncthreads();
    } else if (exact_only || bad) {
      // hand the work item to the exact kernel
      if (threadIdx.x == 0) p.fb_list[atomicAdd(p.fb_count, 1u)] = w;
      __syncthreads();
    } else {
      st.plo = plo;
      PfThresh th;
      th.nhi2 = pk2(-hi, -hi);
      th.band = (int32_t)__float_as_uint(-__fsub_ru(hi, lo));
      for (uint32_t c = c0 + warp; c < c1;) {
        const CellRuns r = s_desc[c - c0];
        pf_cell<CMP>(r, st, th, c2, cons, qn);
        uint32_t nxt = 0;
        if (lane == 0) nxt = atomicAdd(&s_next, 1u);
        c = __shfl_sync(0xffffffffu, nxt, 0);
      }
      pf_flush<CMP>(st, c2, cons, qn);
      cons.template tile_end<CMP>(w);  // ends with __syncthreads(): the stage may be overwritten
    }
    w = s_tile[par];
    par ^= 1u;
  }
  cons.finish();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicInc(p.tile_done, gridDim.x - 1) == gridDim.x - 1) *p.tile_next = 0u;
  }
}

}  // namespace zb
