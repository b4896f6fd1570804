// Microbenchmark: tcgen05.ld / tcgen05.st throughput (32 lanes x 32 columns x 4 B = 4 KB per warp instruction).
#include <cstdio>
#include "tc_common.cuh"
using namespace lsh;

template <int MODE>   // 0: ld x32 back-to-back (wait every 4), 1: ld x32 + wait each, 2: st x16 stream
__global__ void __launch_bounds__(256, 1) bench(long long *out, int iters, int nwarps) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t t = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      if (MODE == 0) {
        uint32_t a[32], b[32], c[32], d[32];
        tmem_ld32(t, a); tmem_ld32(t + 32, b); tmem_ld32(t + 64, c); tmem_ld32(t + 96, d);
        tmem_ld_wait();
        acc += a[i & 31] + b[(i + 1) & 31] + c[3] + d[5];
      } else if (MODE == 1) {
        uint32_t a[32];
        tmem_ld32(t + (i & 7) * 32, a);
        tmem_ld_wait();
        acc += a[i & 31];
      } else {
        uint32_t a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = acc + k;
        tmem_st16(t + (i & 7) * 16, a);
        if ((i & 3) == 3) tmem_st_wait();
        acc += i;
      }
    }
    tmem_st_wait();
  }
  long long t1 = clock64();
  if (lane == 0 && warp < nwarps) out[warp] = (t1 - t0) + (acc == 0xdeadbeef);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long *d; cudaMalloc(&d, 64);
  long long h[8];
  const int iters = 2048;
  for (int nw : {1, 4, 8}) {
    bench<0><<<1, 256>>>(d, iters, nw); cudaDeviceSynchronize(); bench<0><<<1, 256>>>(d, iters, nw); cudaDeviceSynchronize();
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("ld x32 (4 in flight), %d warps: %.1f cycles per 4 KB load per warp -> %.0f B/clk/SM\n", nw, h[0] / (4.0 * iters), nw * 4096.0 * 4 * iters / h[0]);
    bench<1><<<1, 256>>>(d, iters, nw); cudaDeviceSynchronize(); bench<1><<<1, 256>>>(d, iters, nw); cudaDeviceSynchronize();
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("ld x32 + wait each,   %d warps: %.1f cycles per load (latency-bound)\n", nw, h[0] / (double)iters);
    bench<2><<<1, 256>>>(d, iters, nw); cudaDeviceSynchronize(); bench<2><<<1, 256>>>(d, iters, nw); cudaDeviceSynchronize();
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("st x16,               %d warps: %.1f cycles per 2 KB store per warp -> %.0f B/clk/SM  (%s)\n", nw, h[0] / (double)iters, nw * 2048.0 * iters / h[0], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
