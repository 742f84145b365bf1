// common.cuh -- shared device-side definitions of the zelll-b200 engine (sm_100a).
//
// Data layout in HBM:
//   * input   : caller's packed [n][N] coordinates (T = f32 | f64), read-only;
//   * table   : uint32[ncells + 1], dense over the (windowed) cell box in x-fastest / z-slowest
//               order -- the same order the reference's strides induce (util.rs:200-212) without
//               its +4 padding; after a rebuild table[c] / table[c+1] are the CSR begin / end of
//               cell c in `sorted`;
//   * sorted  : Rec<T>[n], one 16 B (f32) or 32 B (f64) record {x, y, z, label} per particle in
//               cell order.  A whole record is one aligned sector write in the scatter and one
//               TMA-bulk-copyable element in the pair kernels.  This is the reference's
//               CellStorage<(usize, [T; 3])> buffer (storage.rs:48-50) with a u32 label.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace zb {

constexpr int kMaxDim = 3;

// ---------------------------------------------------------------------------------------------
// cell-sorted particle record
template <class T>
struct Rec;

template <>
struct __align__(16) Rec<float> {
  float x, y, z;
  uint32_t label;
};

template <>
struct __align__(32) Rec<double> {
  double x, y, z;
  uint32_t label;
  uint32_t pad;
};

static_assert(sizeof(Rec<float>) == 16, "Rec<float> must be one 16 B vector");
static_assert(sizeof(Rec<double>) == 32, "Rec<double> must be one 32 B sector");

__device__ __forceinline__ Rec<float> load_rec(const Rec<float>* p) {
  float4 v = *reinterpret_cast<const float4*>(p);
  Rec<float> r;
  r.x = v.x; r.y = v.y; r.z = v.z; r.label = __float_as_uint(v.w);
  return r;
}
__device__ __forceinline__ Rec<double> load_rec(const Rec<double>* p) {
  const double2* q = reinterpret_cast<const double2*>(p);
  double2 a = q[0];
  double2 b = q[1];
  Rec<double> r;
  r.x = a.x; r.y = a.y; r.z = b.x;
  r.label = (uint32_t)(__double_as_longlong(b.y) & 0xffffffffll);
  r.pad = 0;
  return r;
}
__device__ __forceinline__ void store_rec(Rec<float>* p, float x, float y, float z, uint32_t label) {
  *reinterpret_cast<float4*>(p) = make_float4(x, y, z, __uint_as_float(label));
}
__device__ __forceinline__ void store_rec(Rec<double>* p, double x, double y, double z, uint32_t label) {
  // ONE 256-bit store (sm_100 STG.256): the record is one aligned 32-byte sector, written whole
  // instead of as two half-sector requests that the L2 has to merge
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(x), "d"(y), "d"(z),
               "d"(__longlong_as_double((long long)label))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// Grid geometry handed to kernels by value.  (inf, cutoff, shape) are the reference's GridInfo
// (util.rs:81-90); the window (wlo, wshape) selects the sub-box of cells this handle stores:
// the whole box on one GPU, a z-slab plus its lower halo layer when sharded.
template <class T>
struct GridParams {
  T inf[kMaxDim];
  T cutoff;
  int shape[kMaxDim];   // reference shape (1 on unused axes)
  int wlo[kMaxDim];     // first stored cell per axis
  int wshape[kMaxDim];  // stored cells per axis
  int ndim;
  uint32_t ncells;      // wshape[0] * wshape[1] * wshape[2]
};

// Rust `as i32` on a float: saturating, NaN -> 0 (util.rs:198, :247, :295).  CUDA's
// float->int conversion has the same semantics, spelled out here for the record.
template <class T>
__device__ __forceinline__ int sat_i32(T v) {
  if (v != v) return 0;
  if (v >= T(2147483648.0)) return 2147483647;
  if (v <= T(-2147483649.0)) return (-2147483647 - 1);
  return (int)v;
}

// cell coordinate along one axis: floor((x - inf) / cutoff) as i32  (util.rs:294-296).
// True IEEE subtraction and division (the TU is compiled with -fmad=false -prec-div=true).
template <class T>
__device__ __forceinline__ int cell_coord(T x, T inf, T cutoff) {
  return sat_i32(floor((x - inf) / cutoff));
}

// local (windowed) dense cell id, or 0xffffffff when the particle is outside the window
template <class T>
__device__ __forceinline__ uint32_t local_cell(const GridParams<T>& g, T x, T y, T z) {
  int cx = cell_coord(x, g.inf[0], g.cutoff) - g.wlo[0];
  int cy = cell_coord(y, g.inf[1], g.cutoff) - g.wlo[1];
  int cz = (g.ndim == 3) ? cell_coord(z, g.inf[2], g.cutoff) - g.wlo[2] : 0;
  bool ok = (unsigned)cx < (unsigned)g.wshape[0] && (unsigned)cy < (unsigned)g.wshape[1] &&
            (unsigned)cz < (unsigned)g.wshape[2];
  if (!ok) return 0xffffffffu;
  return (uint32_t)cx + (uint32_t)g.wshape[0] * ((uint32_t)cy + (uint32_t)g.wshape[1] * (uint32_t)cz);
}

// reference flat key (flat_cell_index, util.rs:291-297) with wrapping i32 arithmetic
template <class T>
__device__ __forceinline__ int ref_key(const GridParams<T>& g, T x, T y, T z) {
  uint32_t s1 = (uint32_t)(g.shape[0] + 4);
  uint32_t s2 = s1 * (uint32_t)(g.shape[1] + 4);
  uint32_t k = (uint32_t)cell_coord(x, g.inf[0], g.cutoff);
  k += (uint32_t)cell_coord(y, g.inf[1], g.cutoff) * s1;
  if (g.ndim == 3) k += (uint32_t)cell_coord(z, g.inf[2], g.cutoff) * s2;
  return (int)k;
}

// packed-input loads: particle i of a [n][NDIM] array
template <class T, int NDIM>
__device__ __forceinline__ void load_point(const T* __restrict__ xyz, uint64_t i, T& x, T& y, T& z) {
  const T* p = xyz + i * NDIM;
  x = __ldg(p);
  y = __ldg(p + 1);
  z = (NDIM == 3) ? __ldg(p + 2) : T(0);
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
template <class T, class Op>
__device__ __forceinline__ T warp_reduce(T v, Op op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace zb
