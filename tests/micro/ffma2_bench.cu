// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2) vs MUFU.EX2 vs broadcast LDS.128 per SM sub-partition.
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2(float a) {
  float d;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(a));
  return d;
}
template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(long long *out, float *sink, int iters) {
  __shared__ float4 sm[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  float a[8];
  uint64_t b[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; b[i] = (uint64_t)__float_as_uint(a[i]) << 32 | __float_as_uint(a[i] + 1); }
  const float m = 0.999f; const uint64_t m2 = (uint64_t)__float_as_uint(m) << 32 | __float_as_uint(m);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = ffma1(a[i], m, a[(i + 1) & 7]);
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = ffma2(b[i], m2, b[(i + 1) & 7]);
    } else if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = ex2(a[i]);
    } else if (MODE == 3) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((uint32_t)__cvta_generic_to_shared(sm) + ((it * 32 + r * 8 + i) & 255) * 16));
          a[i] += v.x;
        }
    } else if (MODE == 4) {   // hash-like mix: 1 broadcast LDS.128 per 2 FFMA2
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          uint64_t lo, hi;
          asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(lo), "=l"(hi) : "r"((uint32_t)__cvta_generic_to_shared(sm) + ((it * 16 + r * 4 + i) & 255) * 16));
          b[i] = ffma2(m2, lo, b[i]);
          b[i + 1] = ffma2(m2, hi, b[i + 1]);
        }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)b[i]);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}
int main() {
  long long *d; cudaMalloc(&d, 64); float *sink; cudaMalloc(&sink, 4096 * 4);
  long long h;
  const int iters = 4096;
  const char *names[] = {"FFMA (3-reg)", "FFMA2", "MUFU.EX2", "LDS.128 broadcast", "LDS.128 + 2 FFMA2"};
  for (int nthreads : {128, 256, 512, 1024}) {
    for (int mode = 0; mode < 5; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (mode) {
          case 0: bench<0><<<1, nthreads>>>(d, sink, iters); break;
          case 1: bench<1><<<1, nthreads>>>(d, sink, iters); break;
          case 2: bench<2><<<1, nthreads>>>(d, sink, iters); break;
          case 3: bench<3><<<1, nthreads>>>(d, sink, iters); break;
          case 4: bench<4><<<1, nthreads>>>(d, sink, iters); break;
        }
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      const double per_warp_inst = (double)h / (iters * 32.0);
      printf("%-20s %4d threads: %.2f cycles per warp-instruction slot per warp, %.2f warp-inst/clk/SM\n", names[mode], nthreads,
             per_warp_inst, (nthreads / 32) / per_warp_inst);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
