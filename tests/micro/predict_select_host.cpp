// Host harness for trax_b200/csrc/predict_select.cuh: runs the two per-thread phases of predict_attend_kernel's key
// selection over all 256 thread ids, with the block barrier between them, on the CPU (g++; no CUDA).  tests/test_predict_host.py
// compares the flags with the reference's priority sort (EA:2073-2084).
#include "../../trax_b200/csrc/predict_select.cuh"

extern "C" void predict_select_host(const int32_t *buckets, const int32_t *qb, int M, int nh, int q_start, int k_sel,
                                    uint8_t *flags) {
  int seg_valid[lsh::PREDICT_THREADS], seg_invalid[lsh::PREDICT_THREADS];
  const lsh::PredictSelect p = {buckets, qb, M, nh, q_start, k_sel};
  for (int tid = 0; tid < lsh::PREDICT_THREADS; ++tid) lsh::predict_select_count(p, tid, flags, seg_valid, seg_invalid);
  // __syncthreads()
  for (int tid = 0; tid < lsh::PREDICT_THREADS; ++tid) lsh::predict_select_rank(p, tid, flags, seg_valid, seg_invalid);
}
