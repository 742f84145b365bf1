// query_kernels.cuh -- point queries and cell inspection.
//
//   query_kernel          CellGrid::query + query_neighbors (cellgrid.rs:360-401): the particles of
//                         the query point's cell and of its Full neighbourhood (iters.rs:44-56),
//                         optionally distance-filtered like the Python binding's `neighbors`
//                         (python/src/lib.rs:229-241).
//   cells_flag/compact    CellGrid::iter() (iters.rs:261-266): the non-empty cells, ascending.
#pragma once

#include "common.cuh"
#include "sparse_kernels.cuh"

namespace zb {

template <class T>
struct QueryParams {
  GridParams<T> g;
  const Rec<T>* sorted;
  const uint32_t* csr;
  T c2;
  int cmp;  // 0 none, 1 <, 2 <=
  const unsigned long long* ukeys;  // sparse grids: compact cell keys (csr = ubegin), else nullptr
  uint32_t nuniq;
};

// One warp per query point.  EMIT = false: counts[q] = number of neighbours, valid[q];
// EMIT = true: labels[offsets[q] ...] filled (order inside a query unspecified, as upstream).
template <class T, bool EMIT, bool SPARSE>
__global__ void __launch_bounds__(256) query_kernel(QueryParams<T> p, const T* __restrict__ queries, uint32_t nq,
                                                    unsigned long long* __restrict__ counts_or_offsets,
                                                    uint8_t* __restrict__ valid, uint32_t* __restrict__ labels) {
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= nq) return;
  const unsigned lane = lane_id();
  const int nd = p.g.ndim;
  T qc[3] = {T(0), T(0), T(0)};
  int ci[3] = {0, 0, 0};
  bool ok = true;
  for (int d = 0; d < nd; ++d) {
    qc[d] = __ldg(queries + (uint64_t)q * nd + d);
    // try_cell_index (util.rs:245-256): -1 <= idx <= shape on every axis
    ci[d] = cell_coord(qc[d], p.g.inf[d], p.g.cutoff);
    ok = ok && ci[d] >= -1 && ci[d] <= p.g.shape[d];
  }
  if (!EMIT) {
    if (lane == 0) valid[q] = ok ? 1 : 0;
  }
  unsigned long long base = EMIT ? counts_or_offsets[q] : 0ull;
  unsigned long long total = 0;
  if (ok) {
    const int z0 = nd == 3 ? -1 : 0, z1 = nd == 3 ? 1 : 0;
    for (int dz = z0; dz <= z1; ++dz)
      for (int dy = -1; dy <= 1; ++dy) {
        const int cy = ci[1] + dy, cz = ci[2] + dz;
        if (cy < 0 || cy >= p.g.shape[1] || cz < 0 || cz >= p.g.shape[2]) continue;
        // the three x-neighbours are consecutive cells: one contiguous record range
        const int xa = max(ci[0] - 1, 0), xb = min(ci[0] + 1, p.g.shape[0] - 1);
        if (xa > xb) continue;
        uint32_t b, e;
        if (SPARSE) {
          const unsigned long long row = (unsigned long long)p.g.shape[0] *
                                         ((unsigned long long)cy + (unsigned long long)p.g.shape[1] * (unsigned long long)cz);
          sparse_range(p.ukeys, p.csr, p.nuniq, row + (unsigned long long)xa, row + (unsigned long long)xb, b, e);
        } else {
          const uint32_t row = (uint32_t)p.g.wshape[0] * ((uint32_t)cy + (uint32_t)p.g.wshape[1] * (uint32_t)cz);
          b = __ldg(p.csr + row + xa);
          e = __ldg(p.csr + row + xb + 1);
        }
        for (uint32_t s = b; s < e; s += 32) {
          const uint32_t k = s + lane;
          bool h = k < e;
          Rec<T> r;
          if (h) {
            r = load_rec(p.sorted + k);
            if (p.cmp != 0) {
              const T dx = qc[0] - r.x, dy2 = qc[1] - r.y, dz2 = qc[2] - r.z;
              const T dsq = (dx * dx + dy2 * dy2) + dz2 * dz2;
              h = p.cmp == 1 ? dsq < p.c2 : dsq <= p.c2;
            }
          }
          const unsigned m = __ballot_sync(0xffffffffu, h);
          if (EMIT && h) labels[base + total + __popc(m & lanemask_lt())] = r.label;
          total += __popc(m);
        }
      }
  }
  if (!EMIT && lane == 0) counts_or_offsets[q] = total;
}

// flags[c] = 1 if cell c is non-empty (then scanned in place by scan_kernel)
__global__ void cells_flag_kernel(const uint32_t* __restrict__ csr, uint32_t ncells, uint32_t* __restrict__ flags) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  flags[c] = csr[c + 1] > csr[c] ? 1u : 0u;
}

// pos = exclusive scan of the flags; writes (reference flat key, begin, count) of every non-empty cell
__global__ void cells_compact_kernel(const uint32_t* __restrict__ csr, const uint32_t* __restrict__ pos, uint32_t ncells,
                                     int shape0, int shape1, int wlo0, int wlo1, int wlo2, int w0, int w1,
                                     int32_t* __restrict__ keys, uint32_t* __restrict__ begin,
                                     uint32_t* __restrict__ count) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const uint32_t b = csr[c], e = csr[c + 1];
  if (e <= b) return;
  const uint32_t o = pos[c];
  const uint32_t cx = c % (uint32_t)w0, r = c / (uint32_t)w0;
  const uint32_t cy = r % (uint32_t)w1, cz = r / (uint32_t)w1;
  const uint32_t s1 = (uint32_t)(shape0 + 4), s2 = s1 * (uint32_t)(shape1 + 4);
  keys[o] = (int32_t)((cx + (uint32_t)wlo0) + (cy + (uint32_t)wlo1) * s1 + (cz + (uint32_t)wlo2) * s2);
  begin[o] = b;
  count[o] = e - b;
}

}  // namespace zb
