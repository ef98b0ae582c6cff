// rcp_probe.cu -- dev probe for the EvaluateM inner term (not part of libahx).
//   accuracy (max relative error of 1/d vs IEEE division, d = 1..2^25) and throughput
//   (SMSP cycles per warp-term) of candidate reciprocal sequences for  acc += n / d:
//     V0  MUFU.RCP64H seed + 2 Newton steps            (4 DFMA)
//     V1  MUFU.RCP64H seed + 1 cubic step              (3 DFMA)
//     V2  fp32 MUFU.RCP seed (bit-converted) + 1 Newton (2 DFMA)
//     V3  fp32 MUFU.RCP seed (bit-converted) + 1 cubic  (3 DFMA)
//     V4  MUFU.RCP64H seed + 1 Newton step              (2 DFMA)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rcp_probe rcp_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double u32_to_double(uint32_t v) { return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0; }
__device__ __forceinline__ double rcp64h(double d) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d)); return r; }
// 1/x seed through the fp32 MUFU: x is a normal double in [1, 2^32]
__device__ __forceinline__ double rcp32(double x) {
  const uint32_t hi = __double2hiint(x), lo = __double2loint(x);
  const uint32_t fb = __funnelshift_l(lo, hi - 0x38000000u, 3);      // truncated fp32 bits of x
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__uint_as_float(fb)));
  const uint32_t rb = __float_as_uint(r);
  return __hiloint2double((rb >> 3) + 0x38000000u, rb << 29);
}
template <int V> __device__ __forceinline__ double recip(double x) {
  double r = (V == 2 || V == 3) ? rcp32(x) : rcp64h(x);
  double e = fma(-x, r, 1.0);
  if (V == 1 || V == 3) { e = fma(e, e, e); return fma(r, e, r); }
  r = fma(r, e, r);
  if (V == 0) { e = fma(-x, r, 1.0); r = fma(r, e, r); }
  return r;
}

template <int V> __global__ void accuracy(unsigned long long *out) {
  double m = 0;
  for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x + 1; d <= (1u << 25); d += gridDim.x * blockDim.x) {
    const double x = u32_to_double(d), ex = 1.0 / x;
    m = fmax(m, fabs(recip<V>(x) - ex) / ex);
  }
  atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

template <int V> __global__ void throughput(double *out, int iters, uint32_t seed) {
  constexpr int C = 16;
  uint32_t d[C]; double acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { d[c] = seed + threadIdx.x * 977u + c * 131u + 1u; acc[c] = 0.0; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const uint32_t dd = d[c];
      const double n = __hiloint2double(dd <= 20000000u ? 0x40340000 : 0, 0);   // 20.0 or 0
      acc[c] = fma(n, recip<V>(u32_to_double(dd)), acc[c]);
      d[c] = __usad(dd, 12345u + i, 0u) + 1u;
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V> void run(unsigned long long *d_out, double *buf, double clk_khz) {
  cudaMemset(d_out, 0, 8);
  accuracy<V><<<148 * 8, 256>>>(d_out);
  unsigned long long h; cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
  double e; memcpy(&e, &h, 8);
  printf("V%d max rel err %.3e (2^%.1f)", V, e, e > 0 ? log2(e) : -99.0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int warps = 8; warps <= 32; warps *= 2) {
    const int iters = 4096; float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(a);
      throughput<V><<<148, warps * 32>>>(buf, iters, 7u);
      cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    }
    printf(" | %2d warps/SM: %.2f Tterm/s %.2f clk/warp-term/SMSP", warps, 148.0 * warps * 32 * 16 * iters / ms / 1e9, ms * 1e-3 * clk_khz * 1e3 / (iters * 16.0 * warps / 4.0));
  }
  printf("\n");
}

int main() {
  unsigned long long *d_out; cudaMalloc(&d_out, 8);
  double *buf; cudaMalloc(&buf, 148 * 1024 * 8);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  run<0>(d_out, buf, clk); run<1>(d_out, buf, clk); run<2>(d_out, buf, clk); run<3>(d_out, buf, clk); run<4>(d_out, buf, clk);
  return 0;
}
