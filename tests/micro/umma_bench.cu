// Microbenchmark (not part of the library): cycles per tcgen05.mma for the operand flavours used by the kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I trax_b200/csrc -o /tmp/umma_bench tests/micro/umma_bench.cu
#include <cstdio>
#include "tc_common.cuh"
using namespace lsh;

template <int MODE, int N, int CHAIN>
__global__ void __launch_bounds__(128, 1) bench(long long *out, int n_mma) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base, sa = smem_u32(smem), sb = sa + 32768;
  if (warp == 1) {
    const uint32_t HI = desc_hi(1024);
    const uint32_t idesc_kk = make_idesc_bf16(128, N, 0, 0), idesc_kmn = make_idesc_bf16(128, N, 0, 1),
                   idesc_mnmn = make_idesc_bf16(128, N, 1, 1);
    const uint32_t alo_k = desc_lo(sa, 16), blo_k = desc_lo(sb, 16), blo_mn = desc_lo(sb, 1024), alo_mn = desc_lo(sa, 16384);
    long long t0 = clock64();
    if (elect_one()) {
      for (int i0 = 0; i0 < n_mma; i0 += 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t d = tmem + (CHAIN ? 0 : (i & 1) * 256);
          const uint32_t k = (i & 3);
          if (MODE == 0) umma_ss2(d, alo_k + k * 2, HI, blo_k + k * 2, HI, idesc_kk, 1);
          if (MODE == 1) umma_ts2(d, tmem + 128 + k * 8, blo_k + k * 2, HI, idesc_kk, 1);
          if (MODE == 2) umma_ts2(d, tmem + 128 + k * 8, blo_mn + k * 128, HI, idesc_kmn, 1);
          if (MODE == 3) umma_ss2(d, alo_mn + k * 128, HI, blo_mn + k * 128, HI, idesc_mnmn, 1);
          if (MODE == 4) umma_ss2(d, alo_k + k * 2, HI, blo_mn + k * 128, HI, idesc_kmn, 1);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long *d; cudaMalloc(&d, 16);
  const char *names[] = {"SS  A K-major, B K-major ", "TS  A tmem,    B K-major ", "TS  A tmem,    B MN-major", "SS  A MN-major,B MN-major", "SS  A K-major, B MN-major"};
#define RUN(MODE, N, CHAIN) { long long h[2]; auto kfn = bench<MODE, N, CHAIN>; \
    cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); \
    for (int rep = 0; rep < 2; ++rep) { kfn<<<1, 128, 66 * 1024>>>(d, 512); cudaDeviceSynchronize(); } \
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); \
    printf("N=%3d %s %s : issue %6.1f cyc/mma, complete %6.1f cyc/mma  (%s)\n", N, CHAIN ? "chained " : "2 accums", names[MODE], h[0] / 512.0, h[1] / 512.0, cudaGetErrorString(cudaGetLastError())); }
  RUN(0, 64, 1) RUN(1, 64, 1) RUN(2, 64, 1) RUN(3, 64, 1) RUN(4, 64, 1)
  RUN(0, 64, 0) RUN(2, 64, 0) RUN(3, 64, 0)
  RUN(0, 128, 1) RUN(1, 128, 1) RUN(2, 128, 1) RUN(3, 128, 1) RUN(0, 128, 0) RUN(0, 256, 1) RUN(2, 256, 1)
  return 0;
}
