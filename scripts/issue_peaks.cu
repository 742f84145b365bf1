// issue_peaks.cu -- instruction-issue microbenchmarks on B200 (sm_100a): the ceilings the pair kernels
// are measured against (SURVEY.md section 8d asks for a measured FP64 figure; MEASURED_PEAKS.json has
// only HBM and bf16 tensor numbers).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o build/issue_peaks scripts/issue_peaks.cu
//   build/issue_peaks > profiles/fp64_peak.json
//
// Every kernel runs `iters` iterations of CHAINS independent dependency chains per thread; the rate
// is reported as warp instructions per cycle per SM (from clock64 of the slowest block) and as lane
// operations per second (CUDA events).  No memory traffic except where stated.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e__ = (x);                                                                        \
    if (e__ != cudaSuccess) {                                                                     \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e__));                                   \
      exit(1);                                                                                    \
    }                                                                                             \
  } while (0)

constexpr int CH = 8;

struct Out {
  unsigned long long cycles;
  double sink;
};

__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

enum Kind { DADD, DMUL, DFMA, DTEST, DDIV, FADD, FFMA, FADD2, FFMA2, F2TEST, F2TEST_LDS, BALLOT, COUNT_ };
static const char* kNames[] = {"dadd", "dmul", "dfma", "f64_distance_test", "f64_div", "fadd", "ffma",
                               "fadd2_packed", "ffma2_packed", "f32x2_distance_test", "f32x2_distance_test_lds",
                               "ballot_popc_compaction"};
// warp instructions of the measured kind per chain iteration, and lane operations per instruction
static const int kInstr[] = {1, 1, 1, 9, 1, 1, 1, 1, 1, 8, 9, 7};
static const int kLaneOps[] = {1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 1};

template <int K>
__global__ void __launch_bounds__(256) bench(int iters, double seed, Out* out) {
  __shared__ float4 s_home[64];
  if (threadIdx.x < 64) s_home[threadIdx.x] = make_float4(threadIdx.x * 0.5f, 1.f, 2.f, 0.f);
  __syncthreads();
  double d[CH];
  float f[CH];
  uint64_t p[CH];
  uint32_t m[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    d[c] = seed + c + threadIdx.x * 1e-3;
    f[c] = (float)d[c];
    p[c] = pack2(f[c], f[c] + 0.5f);
    m[c] = 0;
  }
  const double a = seed * 1.0000001, b = seed * 0.25;
  const float fa = (float)a, fb = (float)b;
  const uint64_t pa = pack2(fa, fa), pb = pack2(fb, fb);
  uint32_t qa = 0;
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      if (K == DADD) d[c] = __dadd_rn(d[c], a);
      if (K == DMUL) d[c] = __dmul_rn(d[c], a);
      if (K == DFMA) d[c] = __fma_rn(d[c], a, b);
      if (K == DTEST) {  // the exact test of the pair kernels: 3 sub, 3 mul, 2 add, 1 compare
        const double dx = __dadd_rn(d[c], -a), dy = __dadd_rn(d[c], -b), dz = __dadd_rn(d[c], a);
        const double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (s < b) m[c] += 1;
        d[c] = __longlong_as_double(__double_as_longlong(d[c]) ^ (it & 1));
      }
      if (K == DDIV) d[c] = 1.0 / d[c];
      if (K == FADD) f[c] = __fadd_rn(f[c], fa);
      if (K == FFMA) f[c] = __fmaf_rn(f[c], fa, fb);
      if (K == FADD2) p[c] = add2(p[c], pa);
      if (K == FFMA2) p[c] = fma2(p[c], pa, pb);
      if (K == F2TEST || K == F2TEST_LDS) {  // 3 FADD2, 3 FFMA2, 2 SHF (+ 1 broadcast LDS.128)
        float4 h = make_float4(fa, fb, fa, 0.f);
        if (K == F2TEST_LDS) h = s_home[(it + c) & 63];
        const uint64_t dx = add2(pack2(h.x, h.x), p[c]), dy = add2(pack2(h.y, h.y), p[c]), dz = add2(pack2(h.z, h.z), p[c]);
        const uint64_t s = fma2(dz, dz, fma2(dy, dy, fma2(dx, dx, pb)));
        m[c] = __funnelshift_l((uint32_t)s, m[c], 1);
        m[(c + 1) % CH] = __funnelshift_l((uint32_t)(s >> 32), m[(c + 1) % CH], 1);
      }
      if (K == BALLOT) {  // the per-test compaction of the round-1 kernels: ballot, 2 popc, 2 mad, predicated store
        const bool h = ((m[c] + it) & 3u) == 0u;
        const unsigned bb = __ballot_sync(0xffffffffu, h);
        if (h) m[c] = qa + __popc(bb & lt) * 8u;
        qa += __popc(bb) * 8u;
      }
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += d[c] + f[c] + (double)(p[c] & 0xffff) + m[c];
  s += qa;
  if (threadIdx.x == 0) {
    out[blockIdx.x].cycles = (unsigned long long)(t1 - t0);
    out[blockIdx.x].sink = s;
  }
}

template <int K>
void run(int sms, int ctas_per_sm, int iters, double clock_ghz, bool last) {
  const int blocks = sms * ctas_per_sm;
  Out* d_out;
  CK(cudaMalloc(&d_out, blocks * sizeof(Out)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  bench<K><<<blocks, 256>>>(iters / 8, 1.5, d_out);  // warm-up
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    bench<K><<<blocks, 256>>>(iters, 1.5, d_out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  std::vector<Out> h(blocks);
  CK(cudaMemcpy(h.data(), d_out, blocks * sizeof(Out), cudaMemcpyDeviceToHost));
  unsigned long long cyc = 0;
  for (auto& o : h) cyc = o.cycles > cyc ? o.cycles : cyc;
  const double warp_instr_per_sm = (double)ctas_per_sm * 8 /*warps*/ * CH * (double)iters * kInstr[K];
  const double lane_ops = warp_instr_per_sm * sms * 32.0 * kLaneOps[K];
  printf("  \"%s\": {\"ctas_per_sm\": %d, \"ms\": %.4f, \"cycles\": %llu, \"warp_instr_per_clk_per_sm\": %.3f, "
         "\"lane_ops_per_s\": %.4e, \"units_per_s\": %.4e, \"sm_ghz_effective\": %.3f}%s\n",
         kNames[K], ctas_per_sm, best, cyc, warp_instr_per_sm / (double)cyc, lane_ops / (best * 1e-3),
         lane_ops / kInstr[K] / (best * 1e-3), (double)cyc / (best * 1e-3) / 1e9, last ? "" : ",");
  (void)clock_ghz;
  CK(cudaFree(d_out));
}

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  const int iters = argc > 1 ? atoi(argv[1]) : 20000;
  const int occ = argc > 2 ? atoi(argv[2]) : 4;  // 256-thread CTAs per SM
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"iters\": %d, \"chains_per_thread\": %d,\n", prop.name, sms, iters, CH);
  printf("  \"note\": \"units_per_s = tests (or single instructions) per second over the whole GPU; "
         "f64_distance_test = 3 DADD + 3 DMUL + 2 DADD + DSETP per lane, the reference's exact test; "
         "f32x2_distance_test = 3 FADD2 + 3 FFMA2 + 2 SHF per TWO lane tests\",\n");
  run<DADD>(sms, occ, iters, 0, false);
  run<DMUL>(sms, occ, iters, 0, false);
  run<DFMA>(sms, occ, iters, 0, false);
  run<DTEST>(sms, occ, iters / 4, 0, false);
  run<DDIV>(sms, occ, iters / 8, 0, false);
  run<FADD>(sms, occ, iters, 0, false);
  run<FFMA>(sms, occ, iters, 0, false);
  run<FADD2>(sms, occ, iters, 0, false);
  run<FFMA2>(sms, occ, iters, 0, false);
  run<F2TEST>(sms, occ, iters / 4, 0, false);
  run<F2TEST_LDS>(sms, occ, iters / 4, 0, false);
  run<BALLOT>(sms, occ, iters / 4, 0, true);
  printf("}\n");
  return 0;
}
