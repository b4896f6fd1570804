// Microbenchmark: do two warps issuing tcgen05.mma concurrently slow each other down?
#include <cstdio>
#include "tc_common.cuh"
using namespace lsh;

template <int TWO>
__global__ void __launch_bounds__(128, 1) bench(long long *out, int n_mma) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base, sa = smem_u32(smem), sb = sa + 32768;
  const uint32_t HI = desc_hi(1024);
  const uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0), idesc_o = make_idesc_bf16(128, 64, 0, 1);
  const uint32_t alo_k = desc_lo(sa, 16), blo_k = desc_lo(sb, 16), blo_mn = desc_lo(sb, 1024);
  long long t0 = clock64();
  if (warp == 1) {          // "S issuer": SS N=128 chains of 4 into accumulator region 0
    if (elect_one()) {
      for (int i0 = 0; i0 < n_mma; i0 += 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) umma_ss2(tmem + (i >> 2) * 128, alo_k + (i & 3) * 2, HI, blo_k + (i & 3) * 2, HI, idesc_s, (i & 3) > 0);
      }
      umma_commit(&bar[0]);
    }
    __syncwarp();
    mbar_wait(&bar[0], 0);
    long long t2 = clock64();
    if (lane == 0) out[0] = t2 - t0;
  }
  if (warp == 2 && TWO) {   // "PV issuer": TS N=64 chain of 16 into accumulator at column 384, A at column 256
    if (elect_one()) {
      for (int i0 = 0; i0 < n_mma * 2; i0 += 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) umma_ts2(tmem + 384, tmem + 256 + (i & 7) * 8, blo_mn + (i & 7) * 128, HI, idesc_o, i > 0);
      }
      umma_commit(&bar[1]);
    }
    __syncwarp();
    mbar_wait(&bar[1], 0);
    long long t2 = clock64();
    if (lane == 0) out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long *d; cudaMalloc(&d, 16);
  long long h[2];
  const int n = 512;
  cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int rep = 0; rep < 2; ++rep) { bench<0><<<1, 128, 66 * 1024>>>(d, n); cudaDeviceSynchronize(); }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("S issuer alone      : %lld cycles for %d SS N=128 MMAs (%.1f/mma; nominal 64)\n", h[0], n, h[0] / (double)n);
  for (int rep = 0; rep < 2; ++rep) { bench<1><<<1, 128, 66 * 1024>>>(d, n); cudaDeviceSynchronize(); }
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("S + PV concurrently : S done %lld, PV done %lld cycles; nominal total %d (%s)\n", h[0], h[1], n * 64 + n * 2 * 32, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
