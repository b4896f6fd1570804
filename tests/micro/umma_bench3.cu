// Microbenchmark: tcgen05.mma in SHORT bursts separated by tcgen05.commit (the attention kernels' issue pattern) against one
// long chain: does a commit, or a switch of accumulator / operand flavour, cost pipeline time?
#include <cstdio>
#include "tc_common.cuh"
using namespace lsh;

// mode 0: SS N=64 K-major, bursts of 4 (chain on one accumulator), no commits until the end
// mode 1: same, tcgen05.commit after every burst (nobody waits)
// mode 2: same, commit after every burst AND the issuer waits for it (round trip)
// mode 3: TS N=64 (A from TMEM, B MN-major), bursts of 4, commit per burst
// mode 4: SS N=32 K-major, bursts of 4, commit per burst
// mode 5: SS N=64, A and B MN-major (the dQ product), bursts of 8, commit per burst
// mode 6: SS N=128 K-major, bursts of 4, commit per burst
// mode 7: mixed like one backward item: 8 SS N=64 + commit, 8 TS + 8 SS N=64 + commit, 8 SS MN/MN + commit
__global__ void __launch_bounds__(128, 1) bench(long long *out, int n_burst, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_base, sa = smem_u32(smem), sb = sa + 32768, sc = sa + 65536;
  const uint32_t HI = desc_hi(1024);
  const uint32_t id64 = make_idesc_bf16(128, 64, 0, 0), id32 = make_idesc_bf16(128, 32, 0, 0), id128 = make_idesc_bf16(128, 128, 0, 0);
  const uint32_t id_kv = make_idesc_bf16(128, 64, 0, 1), id_dq = make_idesc_bf16(128, 64, 1, 1);
  const uint32_t a_k = desc_lo(sa, 16), b_k = desc_lo(sb, 16), b_mn = desc_lo(sb, 1024), a_mn = desc_lo(sc, 16384);
  if (warp == 1) {
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int i0 = 0; i0 < n_burst; ++i0) {
      const uint32_t acc = tmem + (i0 & 1) * 128;
      if (elect_one()) {
        if (mode <= 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) umma_ss2(acc, a_k + i * 2, HI, b_k + i * 2, HI, id64, i > 0);
        } else if (mode == 3) {
#pragma unroll
          for (int i = 0; i < 4; ++i) umma_ts2(tmem + 384, tmem + 256 + i * 8, b_mn + i * 128, HI, id_kv, i > 0);
        } else if (mode == 4) {
#pragma unroll
          for (int i = 0; i < 4; ++i) umma_ss2(acc, a_k + i * 2, HI, b_k + i * 2, HI, id32, i > 0);
        } else if (mode == 5) {
#pragma unroll
          for (int i = 0; i < 8; ++i) umma_ss2(tmem + 384, a_mn + i * 128, HI, b_mn + i * 128, HI, id_dq, i > 0);
        } else if (mode == 6) {
#pragma unroll
          for (int i = 0; i < 4; ++i) umma_ss2(acc, a_k + i * 2, HI, b_k + i * 2, HI, id128, i > 0);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) umma_ss2(tmem + (i >> 2) * 64, a_k + (i & 3) * 2, HI, b_k + (i & 3) * 2, HI, id64, (i & 3) > 0);
          umma_commit(&bar[1]);
#pragma unroll
          for (int i = 0; i < 8; ++i) umma_ts2(tmem + 320 - (i >> 2) * 64, tmem + (i >> 2) * 64 + (i & 3) * 8, b_mn + (i & 3) * 128, HI, id_kv, 1);
#pragma unroll
          for (int i = 0; i < 8; ++i) umma_ss2(tmem + 128 + (i >> 2) * 64, a_k + (i & 3) * 2, HI, b_k + (i & 3) * 2, HI, id64, (i & 3) > 0);
          umma_commit(&bar[1]);
#pragma unroll
          for (int i = 0; i < 8; ++i) umma_ss2(tmem + 384, a_mn + i * 128, HI, b_mn + i * 128, HI, id_dq, i > 0);
        }
        if (mode >= 1) umma_commit(&bar[0]);
      }
      __syncwarp();
      if (mode == 2) { mbar_wait(&bar[0], ph); ph ^= 1; }
    }
    long long t1 = clock64();
    if (elect_one()) umma_commit(&bar[1]);
    __syncwarp();
    if (mode != 2) { /* drain: wait for the last commit on bar[1] — its phase flips once per commit, poll until quiet */ }
    // simple drain: spin a fixed time, then stamp completion via a final commit on a fresh barrier is not possible here;
    // use tcgen05 fence + wait on bar[1]'s current phase parity instead
    long long t2 = t1;
    for (int spin = 0; spin < 2000000; ++spin) {
      // done when no phase flip of bar[1] / bar[0] is observed for a while: approximated by waiting on elapsed time
      t2 = clock64();
      if (t2 - t1 > 200000) break;
    }
    if (lane == 0) { out[0] = t1 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long *d; cudaMalloc(&d, 16);
  long long h[2];
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char *names[] = {"SS N64 x4, no commit", "SS N64 x4 + commit", "SS N64 x4 + commit + wait", "TS N64 x4 + commit", "SS N32 x4 + commit",
                         "SS MN/MN N64 x8 + commit", "SS N128 x4 + commit", "mixed bwd item (32 MMAs, 3 commits)"};
  const int per[] = {4, 4, 4, 4, 4, 8, 4, 32};
  for (int mode = 0; mode < 8; ++mode) {
    const int n = 2000;
    for (int rep = 0; rep < 2; ++rep) { bench<<<1, 128, 98 * 1024>>>(d, n, mode); cudaDeviceSynchronize(); }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-40s issue time %8lld cycles / %d bursts = %7.1f per burst = %6.1f per MMA   (%s)\n", names[mode], h[0], n, double(h[0]) / n,
           double(h[0]) / n / per[mode], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
