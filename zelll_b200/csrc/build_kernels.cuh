// build_kernels.cuh -- grid construction on the device (CellGrid::new / rebuild / rebuild_mut).
//
//   K1 bbox_kernel     Aabb::from_particles        (util.rs:35-52)
//   K2 count_kernel    flat_cell_index + counting   (util.rs:291-297, cellgrid.rs:196-204)
//   K3 scan_kernel     reserve_cell slice layout    (cellgrid.rs:207-209, storage.rs:106-111)
//   K4 scatter_kernel  CellStorage::push            (cellgrid.rs:215-231, storage.rs:77-81)
//
// The reference's HashMap<i32, CellSliceMeta> is replaced by a dense uint32 table over the cell
// box: counting uses one L2 atomic per particle, the single-pass decoupled look-back scan turns
// counts into CSR offsets in place, and the scatter's fetch-add on the same table both ranks the
// particle inside its cell and leaves table[c] = end of cell c.
#pragma once

#include "common.cuh"

namespace zb {

// ---------------------------------------------------------------------------------------------
// K1: bounding box.  The packed [n][NDIM] array is read as a flat stream of 16-byte vectors
// (fully coalesced); the grid stride is a multiple of NDIM vectors so every register slot of a
// thread always sees the same axis.
constexpr int kBboxThreads = 384;  // multiple of 2 and 3

template <class T>
struct Vec16;
template <>
struct Vec16<float> {
  using type = float4;
  static constexpr int n = 4;
};
template <>
struct Vec16<double> {
  using type = double2;
  static constexpr int n = 2;
};

template <class T>
__device__ __forceinline__ void vec_unpack(const float4& v, T* e) {
  e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
}
template <class T>
__device__ __forceinline__ void vec_unpack(const double2& v, T* e) {
  e[0] = v.x; e[1] = v.y;
}

// Streaming load that does not allocate in L1 (each input byte is used once per pass).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

template <class T>
__device__ __forceinline__ T pos_inf();
template <>
__device__ __forceinline__ float pos_inf<float>() { return __int_as_float(0x7f800000); }
template <>
__device__ __forceinline__ double pos_inf<double>() { return __longlong_as_double(0x7ff0000000000000ll); }

// partials: [gridDim.x][6] (min xyz, max xyz); out6: final (min xyz, max xyz).
// VEC = elements per load: Vec16<T>::n when xyz is 16 B aligned, else 1.
template <class T, int NDIM, int VEC>
__global__ void __launch_bounds__(kBboxThreads) bbox_kernel(const T* __restrict__ xyz, uint64_t n,
                                                            T* __restrict__ partials,
                                                            unsigned* __restrict__ ticket,
                                                            T* __restrict__ out6) {
  const uint64_t total = n * NDIM;
  const uint64_t nv = total / VEC;
  const uint64_t S = (uint64_t)gridDim.x * blockDim.x;  // multiple of NDIM
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;

  T mn[VEC], mx[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) { mn[e] = pos_inf<T>(); mx[e] = -pos_inf<T>(); }

  if constexpr (VEC > 1) {
    using V = typename Vec16<T>::type;
    const V* xv = reinterpret_cast<const V*>(xyz);
    uint64_t v = t;
    // 4 independent loads in flight per thread
    for (; v + 3 * S < nv; v += 4 * S) {
      V a = ld_stream(xv + v), b = ld_stream(xv + v + S), c = ld_stream(xv + v + 2 * S),
        d = ld_stream(xv + v + 3 * S);
      T ea[VEC], eb[VEC], ec[VEC], ed[VEC];
      vec_unpack<T>(a, ea); vec_unpack<T>(b, eb); vec_unpack<T>(c, ec); vec_unpack<T>(d, ed);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        T lo = ea[e] < eb[e] ? ea[e] : eb[e];
        T lo2 = ec[e] < ed[e] ? ec[e] : ed[e];
        T hi = ea[e] > eb[e] ? ea[e] : eb[e];
        T hi2 = ec[e] > ed[e] ? ec[e] : ed[e];
        lo = lo < lo2 ? lo : lo2;
        hi = hi > hi2 ? hi : hi2;
        mn[e] = lo < mn[e] ? lo : mn[e];
        mx[e] = hi > mx[e] ? hi : mx[e];
      }
    }
    for (; v < nv; v += S) {
      V a = ld_stream(xv + v);
      T ea[VEC];
      vec_unpack<T>(a, ea);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        mn[e] = ea[e] < mn[e] ? ea[e] : mn[e];
        mx[e] = ea[e] > mx[e] ? ea[e] : mx[e];
      }
    }
  } else {
    for (uint64_t v = t; v < nv; v += S) {
      T a = __ldg(xyz + v);
      mn[0] = a < mn[0] ? a : mn[0];
      mx[0] = a > mx[0] ? a : mx[0];
    }
  }

  // fold the per-slot extrema onto axes: slot e of this thread always held axis (t*VEC+e) % NDIM
  T cm[3] = {pos_inf<T>(), pos_inf<T>(), pos_inf<T>()};
  T cM[3] = {-pos_inf<T>(), -pos_inf<T>(), -pos_inf<T>()};
#pragma unroll
  for (int e = 0; e < VEC; ++e) {
    int ax = (int)((t * VEC + e) % NDIM);
#pragma unroll
    for (int d = 0; d < NDIM; ++d)
      if (ax == d) {
        cm[d] = mn[e] < cm[d] ? mn[e] : cm[d];
        cM[d] = mx[e] > cM[d] ? mx[e] : cM[d];
      }
  }
  // scalar tail (total % VEC elements), one element per thread of block 0
  if (VEC > 1 && blockIdx.x == 0) {
    uint64_t s = nv * VEC + threadIdx.x;
    if (s < total) {
      T a = __ldg(xyz + s);
      int ax = (int)(s % NDIM);
#pragma unroll
      for (int d = 0; d < NDIM; ++d)
        if (ax == d) {
          cm[d] = a < cm[d] ? a : cm[d];
          cM[d] = a > cM[d] ? a : cM[d];
        }
    }
  }

  auto fmin_ = [](T a, T b) { return a < b ? a : b; };
  auto fmax_ = [](T a, T b) { return a > b ? a : b; };
  __shared__ T s_red[kBboxThreads / 32][6];
  __shared__ bool s_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    T a = warp_reduce(cm[d], fmin_);
    T b = warp_reduce(cM[d], fmax_);
    if (lane == 0) { s_red[warp][d] = a; s_red[warp][3 + d] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    T a = s_red[0][threadIdx.x];
    for (int w = 1; w < kBboxThreads / 32; ++w)
      a = threadIdx.x < 3 ? fmin_(a, s_red[w][threadIdx.x]) : fmax_(a, s_red[w][threadIdx.x]);
    partials[(uint64_t)blockIdx.x * 6 + threadIdx.x] = a;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  // last block: fold all partials (fixed order -> deterministic)
  __threadfence();
  T fm[3] = {pos_inf<T>(), pos_inf<T>(), pos_inf<T>()};
  T fM[3] = {-pos_inf<T>(), -pos_inf<T>(), -pos_inf<T>()};
  for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
    const volatile T* p = partials + (uint64_t)b * 6;
#pragma unroll
    for (int d = 0; d < 3; ++d) { fm[d] = fmin_(fm[d], p[d]); fM[d] = fmax_(fM[d], p[3 + d]); }
  }
  __syncthreads();
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    T a = warp_reduce(fm[d], fmin_);
    T b = warp_reduce(fM[d], fmax_);
    if (lane == 0) { s_red[warp][d] = a; s_red[warp][3 + d] = b; }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    T a = s_red[0][threadIdx.x];
    for (int w = 1; w < kBboxThreads / 32; ++w)
      a = threadIdx.x < 3 ? fmin_(a, s_red[w][threadIdx.x]) : fmax_(a, s_red[w][threadIdx.x]);
    out6[threadIdx.x] = a;
    if (threadIdx.x == 0) *ticket = 0;  // re-arm for the next rebuild
  }
}

// ---------------------------------------------------------------------------------------------
// K2: per-particle cell id + histogram.  counts[c] += 1 with one no-return L2 atomic (RED).
constexpr int kPointThreads = 256;

// Each thread handles kPointIlp particles (block-strided so loads stay coalesced): all loads are
// issued before the first atomic, which keeps kPointIlp L2 round trips in flight per thread.
constexpr int kPointIlp = 1;  // measured: 4 is ~10 % slower (the kernels are bound by random sector traffic, not latency)

// particle labels travel through f32 / f64 halo rows as raw bits
template <class T>
__device__ __forceinline__ T label_bits(uint32_t label);
template <>
__device__ __forceinline__ float label_bits<float>(uint32_t label) { return __uint_as_float(label); }
template <>
__device__ __forceinline__ double label_bits<double>(uint32_t label) { return __longlong_as_double((long long)label); }
__device__ __forceinline__ uint32_t label_from_bits(float v) { return __float_as_uint(v); }
__device__ __forceinline__ uint32_t label_from_bits(double v) { return (uint32_t)__double_as_longlong(v); }

// Slab-local multi-GPU step: while counting its own rows a rank also picks out the particles of its
// TOP layer -- the lower halo of the next rank -- as rows {x, y, z, label} of out[1..] (order
// unspecified; their number accumulates in *count), so the halo costs no extra pass over the input.
// A row outside the rank's own layers [first, top] of the slab axis sets *bad.
template <class T>
struct TopLayerOut {
  T* out;
  uint32_t cap;
  uint32_t* count;
  int* bad;
  uint32_t label_offset;
  int first, top;  // window-relative layers of z_begin and z_end - 1
};

template <class T, int NDIM, bool TOP>
__global__ void __launch_bounds__(kPointThreads) count_kernel(const T* __restrict__ xyz, uint32_t n,
                                                              GridParams<T> g,
                                                              uint32_t* __restrict__ counts,
                                                              int* __restrict__ flags, TopLayerOut<T> tl,
                                                              const uint32_t* __restrict__ n_dev = nullptr) {
  // n_dev (halo rows of the slab-local step): the number of valid rows lives on the device; n is the
  // capacity the grid was sized for
  if (n_dev) n = min(n, *n_dev);
  const uint32_t base = blockIdx.x * (kPointThreads * kPointIlp) + threadIdx.x;
  uint32_t c[kPointIlp];
#pragma unroll
  for (int k = 0; k < kPointIlp; ++k) {
    const uint32_t i = base + k * kPointThreads;
    c[k] = 0xfffffffeu;  // beyond n
    T x = T(0), y = T(0), z = T(0);
    bool top = false;
    if (i < n) {
      load_point<T, NDIM>(xyz, i, x, y, z);
      c[k] = local_cell(g, x, y, z);
      if (TOP) {
        const int layer = cell_coord(NDIM == 3 ? z : y, g.inf[NDIM - 1], g.cutoff) - g.wlo[NDIM - 1];
        if (layer < tl.first || layer > tl.top) atomicOr(tl.bad, 1);
        top = layer == tl.top;
      }
    }
    if (TOP) {  // warp-aggregated append (every thread of the warp gets here)
      const unsigned b = __ballot_sync(0xffffffffu, top);
      if (b != 0) {
        uint32_t slot0 = 0;
        if (lane_id() == 0) slot0 = atomicAdd(tl.count, (uint32_t)__popc(b));
        slot0 = __shfl_sync(0xffffffffu, slot0, 0);
        const uint32_t slot = slot0 + __popc(b & lanemask_lt());
        if (top && slot < tl.cap) {
          T* row = tl.out + (uint64_t)(slot + 1) * 4;
          row[0] = x; row[1] = y; row[2] = z; row[3] = label_bits<T>(tl.label_offset + i);
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kPointIlp; ++k) {
    if (c[k] == 0xffffffffu) atomicOr(flags, 1);
    else if (c[k] != 0xfffffffeu) atomicAdd(counts + c[k], 1u);
  }
}

// ---------------------------------------------------------------------------------------------
// K3: in-place exclusive scan of the count table, single pass with decoupled look-back.
// state[tile] = (status << 32) | value ; status 0 = empty, 1 = tile aggregate, 2 = inclusive prefix.
constexpr int kScanThreads = 512;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kScanThreads) scan_kernel(uint32_t* __restrict__ a, uint32_t m,
                                                            unsigned long long* __restrict__ state,
                                                            uint32_t* __restrict__ tile_counter,
                                                            uint32_t* __restrict__ nonempty) {
  __shared__ uint32_t s_tile;
  __shared__ uint32_t s_warp[kScanThreads / 32];
  __shared__ uint32_t s_prefix;
  __shared__ uint32_t s_nonempty;
  if (threadIdx.x == 0) {
    s_tile = atomicAdd(tile_counter, 1u);
    s_nonempty = 0;
  }
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * kScanTile + threadIdx.x * kScanItems;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  uint32_t v[kScanItems];
  if (base + kScanItems <= m) {
    const uint4* p = reinterpret_cast<const uint4*>(a + base);  // a is 16 B aligned, base % 16 == 0
#pragma unroll
    for (int k = 0; k < kScanItems / 4; ++k) {
      uint4 q = p[k];
      v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) v[k] = (base + k < m) ? a[base + k] : 0u;
  }
  uint32_t sum = 0, nz = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { sum += v[k]; nz += (v[k] != 0u); }

  // block-wide exclusive scan of the thread sums
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  nz = warp_reduce(nz, [](uint32_t x, uint32_t y) { return x + y; });
  if (lane == 31) s_warp[warp] = incl;
  if (lane == 0 && nz) atomicAdd(&s_nonempty, nz);
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < kScanThreads / 32) ? s_warp[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += y;
    }
    if (lane < kScanThreads / 32) s_warp[lane] = wi - w;  // exclusive warp offsets
    uint32_t aggregate = __shfl_sync(0xffffffffu, wi, kScanThreads / 32 - 1);

    // decoupled look-back (warp 0)
    uint32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) st_state(state, (2ull << 32) | aggregate);
    } else {
      if (lane == 0) st_state(state + tile, (1ull << 32) | aggregate);
      int look = (int)tile - 1;
      while (true) {
        int idx = look - lane;
        unsigned long long s = (idx >= 0) ? ld_state(state + idx) : (2ull << 32);
        while (__any_sync(0xffffffffu, (s >> 32) == 0ull)) {
          if ((s >> 32) == 0ull) s = ld_state(state + idx);
        }
        unsigned pm = __ballot_sync(0xffffffffu, (s >> 32) == 2ull);
        int first = pm ? (__ffs(pm) - 1) : 32;
        uint32_t val = (lane <= first) ? (uint32_t)(s & 0xffffffffull) : 0u;
        val = warp_reduce(val, [](uint32_t x, uint32_t y) { return x + y; });
        excl += val;
        if (pm) break;
        look -= 32;
      }
      if (lane == 0) st_state(state + tile, (2ull << 32) | (unsigned long long)(excl + aggregate));
    }
    if (lane == 0) s_prefix = excl;
  }
  __syncthreads();
  uint32_t run = s_prefix + s_warp[warp] + (incl - sum);
  if (base + kScanItems <= m) {
    uint4* p = reinterpret_cast<uint4*>(a + base);
#pragma unroll
    for (int k = 0; k < kScanItems / 4; ++k) {
      uint4 q;
      q.x = run; run += v[4 * k];
      q.y = run; run += v[4 * k + 1];
      q.z = run; run += v[4 * k + 2];
      q.w = run; run += v[4 * k + 3];
      p[k] = q;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (base + k < m) a[base + k] = run;
      run += v[k];
    }
  }
  if (threadIdx.x == 0 && s_nonempty) atomicAdd(nonempty, s_nonempty);
}

// ---------------------------------------------------------------------------------------------
// K4: scatter.  pos = fetch-add(table[c]) ranks the particle inside its cell; the record goes
// out as one aligned 16 B / 32 B write.  After this kernel table[c] == end of cell c.
// where a particle's label comes from: an explicit array, or (slab-local step) label_offset + i for
// the first n_local particles and halo_labels[i - n_local] for the halo rows behind them
struct LabelSrc {
  const uint32_t* labels;
  const uint32_t* halo_labels;
  uint32_t offset, n_local;
  __device__ __forceinline__ uint32_t at(uint32_t i) const {
    if (labels) return __ldg(labels + i);
    if (i < n_local) return offset + i;
    return __ldg(halo_labels + (i - n_local));
  }
};

template <class T, int NDIM>
__global__ void __launch_bounds__(kPointThreads) scatter_kernel(const T* __restrict__ xyz, LabelSrc labels,
                                                                uint32_t n, GridParams<T> g,
                                                                uint32_t* __restrict__ cursor,
                                                                Rec<T>* __restrict__ sorted, uint32_t n_fixed = 0,
                                                                const uint32_t* __restrict__ n_dev = nullptr) {
  // slab-local step: rows [0, n_fixed) are the rank's own, *n_dev halo rows follow (n = capacity)
  if (n_dev) n = min(n, n_fixed + *n_dev);
  const uint32_t base = blockIdx.x * (kPointThreads * kPointIlp) + threadIdx.x;
  T x[kPointIlp], y[kPointIlp], z[kPointIlp];
  uint32_t c[kPointIlp], lab[kPointIlp], pos[kPointIlp];
#pragma unroll
  for (int k = 0; k < kPointIlp; ++k) {
    const uint32_t i = base + k * kPointThreads;
    c[k] = 0xffffffffu;
    lab[k] = i;
    if (i < n) {
      load_point<T, NDIM>(xyz, i, x[k], y[k], z[k]);
      c[k] = local_cell(g, x[k], y[k], z[k]);  // 0xffffffff: outside the window, flagged by count_kernel
      lab[k] = labels.at(i);
    }
  }
#pragma unroll
  for (int k = 0; k < kPointIlp; ++k) pos[k] = c[k] != 0xffffffffu ? atomicAdd(cursor + c[k], 1u) : 0u;
#pragma unroll
  for (int k = 0; k < kPointIlp; ++k)
    if (c[k] != 0xffffffffu) store_rec(sorted + pos[k], x[k], y[k], z[k], lab[k]);
}

// ---------------------------------------------------------------------------------------------
// Optional stable order (zb_grid_set_stable): K4 ranks the particles of a cell by atomic arrival,
// the reference's CellStorage::push keeps input order (storage.rs:77-81).  One thread per cell
// sorts its records by label in place (heap sort: O(m log m), no extra memory); with labels =
// enumerate order this reproduces the reference's within-cell order and makes cell_storage(), the
// pair list order inside a cell and the floating-point summation order run-to-run reproducible.
template <class T>
__device__ __forceinline__ void sift_down(Rec<T>* a, uint32_t start, uint32_t end) {
  uint32_t root = start;
  while (2 * root + 1 < end) {
    uint32_t child = 2 * root + 1;
    if (child + 1 < end && load_rec(a + child).label < load_rec(a + child + 1).label) ++child;
    const Rec<T> r = load_rec(a + root), c = load_rec(a + child);
    if (r.label >= c.label) return;
    store_rec(a + root, c.x, c.y, c.z, c.label);
    store_rec(a + child, r.x, r.y, r.z, r.label);
    root = child;
  }
}

template <class T>
__global__ void cell_sort_kernel(const uint32_t* __restrict__ csr, uint32_t ncells, Rec<T>* __restrict__ sorted) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const uint32_t b = csr[c], m = csr[c + 1] - b;
  if (m < 2) return;
  Rec<T>* a = sorted + b;
  if (m <= 16) {  // insertion sort: the common case (~10 particles per cell)
    for (uint32_t i = 1; i < m; ++i) {
      const Rec<T> key = load_rec(a + i);
      uint32_t j = i;
      while (j > 0) {
        const Rec<T> p = load_rec(a + j - 1);
        if (p.label <= key.label) break;
        store_rec(a + j, p.x, p.y, p.z, p.label);
        --j;
      }
      store_rec(a + j, key.x, key.y, key.z, key.label);
    }
    return;
  }
  for (uint32_t start = m / 2; start-- > 0;) sift_down(a, start, m);
  for (uint32_t end = m - 1; end > 0; --end) {
    const Rec<T> top = load_rec(a), last = load_rec(a + end);
    store_rec(a, last.x, last.y, last.z, last.label);
    store_rec(a + end, top.x, top.y, top.z, top.label);
    sift_down(a, 0u, end);
  }
}

// ---------------------------------------------------------------------------------------------
// inspection helpers

// FlatIndex.index in input order, recomputed from the cell-sorted records (same arithmetic as K2)
template <class T>
__global__ void keys_kernel(const Rec<T>* __restrict__ sorted, uint32_t n, GridParams<T> g,
                            int32_t* __restrict__ out) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  Rec<T> r = load_rec(sorted + p);
  out[r.label] = ref_key(g, r.x, r.y, r.z);
}

// FlatIndex::rebuild_mut's change detection (flatindex.rs:140-152): old keys beyond the old
// length count as 0 (Vec::resize(size, 0), flatindex.rs:130)
__global__ void keys_changed_kernel(const int32_t* __restrict__ old_keys, uint32_t n_old,
                                    const int32_t* __restrict__ new_keys, uint32_t n,
                                    int* __restrict__ changed) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t o = i < n_old ? old_keys[i] : 0;
  if (o != new_keys[i]) *changed = 1;
}

// bbox result (T) -> 6 doubles in device memory, for all-reducing the box without a host round trip
template <class T>
__global__ void widen6_kernel(const T* __restrict__ in6, double* __restrict__ out6, int ndim, int negate_inf) {
  const int k = threadIdx.x;
  if (k < 6) {
    double v = (k % 3) < ndim ? (double)in6[k] : 0.0;
    if (negate_inf && k < 3) v = -v;  // all-reduce(max) of (-inf, sup) = (-global inf, global sup)
    out6[k] = v;
  }
}

// slab-local step run with LAST step's bounding box: does the all-reduced box (-inf, sup as doubles) still
// equal it?  Bit 2 of *flag = no (every rank sees the same box, hence the same verdict).
struct Box6 {
  double v[6];
};
__global__ void spec_check_kernel(const double* __restrict__ red6, Box6 expect, uint32_t* __restrict__ flag) {
  const int k = threadIdx.x;
  if (k < 6 && !(red6[k] == expect.v[k])) atomicOr(flag, 4u);
}

// received halo block -> packed coordinates behind the local particles + their labels
// (the launch is sized for cap rows; the received count stays on the device: *n_out = rows that were taken,
// bit 1 of *flag = the sender's layer or the room behind the local rows was too small)
template <class T, int NDIM>
__global__ void halo_unpack_kernel(const T* __restrict__ rows, uint32_t cap, T* __restrict__ xyz_tail,
                                   uint32_t* __restrict__ halo_labels, uint32_t* __restrict__ n_out = nullptr,
                                   uint32_t* __restrict__ flag = nullptr) {
  const uint32_t n = (uint32_t)rows[0];
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0 && n_out) {
    *n_out = min(n, cap);
    if (n > cap) atomicOr(flag, 2u);
  }
  if (r >= n || r >= cap) return;
  const T* row = rows + (uint64_t)(r + 1) * 4;
  xyz_tail[(uint64_t)r * NDIM] = row[0];
  xyz_tail[(uint64_t)r * NDIM + 1] = row[1];
  if (NDIM == 3) xyz_tail[(uint64_t)r * NDIM + 2] = row[2];
  halo_labels[r] = label_from_bits(row[3]);
}

// halo block header: row 0, value 0 = number of rows that follow (as a number, not as bits)
template <class T>
__global__ void halo_header_kernel(const uint32_t* __restrict__ count, uint32_t cap, T* __restrict__ rows) {
  rows[0] = (T)min(*count, cap + 1u);  // cap + 1 signals overflow to the receiver
}

// per-axis cell coordinate of packed input particles (slab assignment of the sharded host)
template <class T>
__global__ void layer_kernel(const T* __restrict__ xyz, uint32_t n, int ndim, int axis, T inf, T cutoff,
                             int32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = cell_coord(__ldg(xyz + (uint64_t)i * ndim + axis), inf, cutoff);
}


// Slab-local input check + halo extraction of the sharded host: every particle must
// lie in layers [z_begin, z_end) of the slab axis (else *bad = 1); the particles of the TOP layer
// z_end - 1 -- the lower halo of the next rank -- are compacted into rows {x, y, z, label} of
// out[1..] (order unspecified), their number accumulates in *count.
template <class T, int NDIM>
__global__ void __launch_bounds__(256) slab_top_kernel(const T* __restrict__ xyz, uint32_t n, T inf, T cutoff,
                                                       int z_begin, int z_end, uint32_t label_offset,
                                                       T* __restrict__ out, uint32_t cap,
                                                       uint32_t* __restrict__ count, int* __restrict__ bad) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool top = false;
  T x = T(0), y = T(0), z = T(0);
  if (i < n) {
    load_point<T, NDIM>(xyz, i, x, y, z);
    const int layer = cell_coord(NDIM == 3 ? z : y, inf, cutoff);
    if (layer < z_begin || layer >= z_end) atomicOr(bad, 1);
    top = layer == z_end - 1;
  }
  const unsigned b = __ballot_sync(0xffffffffu, top);
  if (b == 0) return;
  uint32_t base = 0;
  if (lane_id() == 0) base = atomicAdd(count, (uint32_t)__popc(b));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (top) {
    const uint32_t slot = base + __popc(b & lanemask_lt());
    if (slot < cap) {
      T* row = out + (uint64_t)(slot + 1) * 4;
      row[0] = x; row[1] = y; row[2] = z; row[3] = label_bits<T>(label_offset + i);
    }
  }
}

// cell_storage(): unpack records into labels / packed coordinates
template <class T, int NDIM>
__global__ void unpack_kernel(const Rec<T>* __restrict__ sorted, uint32_t n, uint32_t* __restrict__ labels,
                              T* __restrict__ xyz) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  Rec<T> r = load_rec(sorted + p);
  if (labels) labels[p] = r.label;
  if (xyz) {
    xyz[(uint64_t)p * NDIM] = r.x;
    xyz[(uint64_t)p * NDIM + 1] = r.y;
    if (NDIM == 3) xyz[(uint64_t)p * NDIM + 2] = r.z;
  }
}

}  // namespace zb
