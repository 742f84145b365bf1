// sparse_kernels.cuh -- grid construction in O(n) memory for boxes with (far) more cells than particles.
//
// The reference keeps its cells in a hash map (`HashMap<i32, CellSliceMeta>`, src/cellgrid.rs:120) and so
// pays only for NON-EMPTY cells (README.md:21-22).  The dense count table of build_kernels.cuh costs 4 bytes
// per cell of the bounding box; when that is out of proportion to n (or beyond 2^31 cells) the rebuild takes
// this path instead: a compact SORTED array of the non-empty cells plus CSR offsets.
//
//   S1 keys64_kernel     cell of every particle as one 64-bit key  cx + W0 (cy + W1 cz)   (x fastest, z
//                        slowest: the order of the dense table, so the z-major half shell of the pair kernels
//                        is again 5 runs of consecutive records)
//   S2 radix sort        (key, particle) pairs by key, least significant digit first, 8 bits per pass, only
//                        the passes the key width needs; hand-written: per-block histograms -> one exclusive
//                        scan (scan_kernel) -> stable per-block scatter (warp match + per-warp digit counts)
//   S3 heads + compact   first record of every distinct key -> ukeys[u], ubegin[u] (u < nuniq), ubegin[nuniq] = n
//   S4 gather_kernel     records into sorted order (stable: particles of a cell keep input order, as
//                        CellStorage::push, storage.rs:77-81)
//
// Neighbour cells are found by binary search in ukeys (cell_runs_sparse in pair_kernels.cuh,
// query_kernel<.., SPARSE>): the three x-neighbours of a row are consecutive keys, hence one record range.
#pragma once

#include "build_kernels.cuh"

namespace zb {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;                          // items per thread
constexpr int kSortTile = kSortThreads * kSortItems;   // items per block

// S1: 64-bit cell keys in input order; flags bit0 = a particle outside the box (non-finite coordinate)
template <class T, int NDIM>
__global__ void __launch_bounds__(256) keys64_kernel(const T* __restrict__ xyz, uint32_t n, GridParams<T> g,
                                                     unsigned long long* __restrict__ keys, uint32_t* __restrict__ idx,
                                                     int* __restrict__ flags) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T x, y, z;
  load_point<T, NDIM>(xyz, i, x, y, z);
  const int cx = cell_coord(x, g.inf[0], g.cutoff), cy = cell_coord(y, g.inf[1], g.cutoff);
  const int cz = NDIM == 3 ? cell_coord(z, g.inf[2], g.cutoff) : 0;
  const bool ok = (unsigned)cx < (unsigned)g.shape[0] && (unsigned)cy < (unsigned)g.shape[1] &&
                  (unsigned)cz < (unsigned)g.shape[2];
  if (!ok) atomicOr(flags, 1);
  // out-of-box particles (NaN coordinates) sort behind every real cell and are dropped by the compaction
  keys[i] = ok ? (unsigned long long)cx +
                     (unsigned long long)g.shape[0] * ((unsigned long long)cy + (unsigned long long)g.shape[1] * (unsigned long long)cz)
               : ~0ull;
  idx[i] = i;
}

// S2a: per-block digit histogram, digit-major layout hist[d * nblk + blk] (one scan gives every block its offsets)
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const unsigned long long* __restrict__ keys, uint32_t n,
                                                                  int shift, uint32_t nblk, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[256];
  sh[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const uint32_t i = base + k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&sh[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * nblk + blockIdx.x] = sh[threadIdx.x];
}

// S2b: stable scatter of one block's items.  offs = the scanned histogram.  Items are taken in rounds of 256
// (round-major, then thread order = input order); inside a round a warp ranks equal digits with match_any,
// warps are ordered through per-warp digit counts.
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const unsigned long long* __restrict__ keys,
                                                                     const uint32_t* __restrict__ idx, uint32_t n, int shift,
                                                                     uint32_t nblk, const uint32_t* __restrict__ offs,
                                                                     unsigned long long* __restrict__ keys_out,
                                                                     uint32_t* __restrict__ idx_out) {
  constexpr int kWarps = kSortThreads / 32;
  __shared__ uint32_t s_base[256];
  __shared__ uint32_t s_cnt[kWarps][256];
  s_base[threadIdx.x] = offs[threadIdx.x * nblk + blockIdx.x];
  const int warp = threadIdx.x >> 5;
  const unsigned lane = lane_id();
  const uint32_t base = blockIdx.x * kSortTile;
  for (int k = 0; k < kSortItems; ++k) {
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t i = base + k * kSortThreads + threadIdx.x;
    const bool live = i < n;
    unsigned long long key = 0;
    uint32_t id = 0, d = 256u + lane;  // dead lanes: a digit nobody shares
    if (live) {
      key = keys[i];
      id = idx[i];
      d = (uint32_t)(key >> shift) & 255u;
    }
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank = __popc(peers & lanemask_lt());
    if (live && rank == 0) s_cnt[warp][d] = __popc(peers);
    __syncthreads();
    if (live) {
      uint32_t before = 0;
      for (int w = 0; w < warp; ++w) before += s_cnt[w][d];
      const uint32_t dst = s_base[d] + before + rank;
      keys_out[dst] = key;
      idx_out[dst] = id;
    }
    __syncthreads();
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) tot += s_cnt[w][threadIdx.x];
    s_base[threadIdx.x] += tot;
    // (the zeroing at the top of the next round is ordered behind these reads by its own barrier)
    __syncthreads();
  }
}

// S3a: flags[p] = 1 at the first record of every distinct (real) key; flags[n] = 0 (its scanned value = nuniq)
__global__ void heads_kernel(const unsigned long long* __restrict__ keys, uint32_t n, uint32_t* __restrict__ flags) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  uint32_t f = 0;
  if (p < n) {
    const unsigned long long k = keys[p];
    f = (k != ~0ull && (p == 0 || keys[p - 1] != k)) ? 1u : 0u;
  }
  flags[p] = f;
}

// S3b: pos = exclusive scan of the flags.  ukeys[u] / ubegin[u] of every distinct key, ubegin[nuniq] = end of
// the last real cell (out-of-box records, if any, sit behind it and belong to no cell)
__global__ void uniq_compact_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ pos, uint32_t n,
                                    unsigned long long* __restrict__ ukeys, uint32_t* __restrict__ ubegin) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const unsigned long long k = keys[p];
  const bool real = k != ~0ull;
  if (real && (p == 0 || keys[p - 1] != k)) {
    ukeys[pos[p]] = k;
    ubegin[pos[p]] = p;
  }
  // end marker: the record after the last real one
  if (real && (p + 1 == n || keys[p + 1] == ~0ull)) ubegin[pos[n]] = p + 1;
}

// S4: records in sorted order
template <class T, int NDIM>
__global__ void __launch_bounds__(256) gather_kernel(const T* __restrict__ xyz, const uint32_t* __restrict__ idx, uint32_t n,
                                                     LabelSrc labels, Rec<T>* __restrict__ sorted) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t i = idx[p];
  T x, y, z;
  load_point<T, NDIM>(xyz, i, x, y, z);
  store_rec(sorted + p, x, y, z, labels.at(i));
}

// CellGrid::iter() over the compact cells: reference flat key (wrapping i32 strides, util.rs:200-212), begin, count
__global__ void cells_sparse_kernel(const unsigned long long* __restrict__ ukeys, const uint32_t* __restrict__ ubegin,
                                    uint32_t nuniq, int shape0, int shape1, int32_t* __restrict__ keys,
                                    uint32_t* __restrict__ begin, uint32_t* __restrict__ count) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nuniq) return;
  const unsigned long long k = ukeys[u];
  const unsigned long long cx = k % (unsigned long long)shape0, r = k / (unsigned long long)shape0;
  const unsigned long long cy = r % (unsigned long long)shape1, cz = r / (unsigned long long)shape1;
  const uint32_t s1 = (uint32_t)(shape0 + 4), s2 = s1 * (uint32_t)(shape1 + 4);
  keys[u] = (int32_t)((uint32_t)cx + (uint32_t)cy * s1 + (uint32_t)cz * s2);
  begin[u] = ubegin[u];
  count[u] = ubegin[u + 1] - ubegin[u];
}

// first index in [lo, hi) with ukeys[index] >= key
__device__ __forceinline__ uint32_t lower_bound_u64(const unsigned long long* __restrict__ a, uint32_t lo, uint32_t hi,
                                                    unsigned long long key) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(a + mid) < key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// records of the cells with keys in [k0, k1] (k1 - k0 <= 2: consecutive keys = consecutive compact cells),
// searched among the compact cells [0, hi)
__device__ __forceinline__ void sparse_range(const unsigned long long* __restrict__ ukeys, const uint32_t* __restrict__ ubegin,
                                             uint32_t hi, unsigned long long k0, unsigned long long k1, uint32_t& b,
                                             uint32_t& e) {
  const uint32_t lo = lower_bound_u64(ukeys, 0u, hi, k0);
  uint32_t up = lo;
  while (up < hi && up < lo + 3u && __ldg(ukeys + up) <= k1) ++up;
  b = __ldg(ubegin + lo);
  e = __ldg(ubegin + up);  // lo == up: empty range (b == e)
}

}  // namespace zb
