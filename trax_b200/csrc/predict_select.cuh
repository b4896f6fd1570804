// Key selection of one fast-inference step of LSHSelfAttention (`_incremental_forward_unbatched`, EA:2056-2084), written as
// per-thread phase functions so that the same code runs inside predict_attend_kernel (predict.cu) and, phase by phase over
// all thread ids, inside the host test harness (tests/micro/predict_select_host.cpp) that checks it against the oracle.
//
// The reference sorts every memory slot i by
//     priority(i) = (i > q_start + 1 ? -(M + i) : i) + M * is_valid_target(i)                      (EA:2078-2081)
// with is_valid_target(i) = "slot i shares a bucket with the query in at least one hash round" (EA:2073), and attends to the
// K = n_hashes * chunk_len * (1 + n_chunks_before) slots of highest priority (EA:2082-2084) under a causal + self mask
// (EA:2091-2092).  Priorities are distinct, so the sort is a ranking: by descending priority come the same-bucket slots
// i <= q_start + 1 from the most recent one back, then the other slots i <= q_start + 1 from the most recent one back, then
// the slots after q_start + 1.  Slots after q_start get probability exp(-1e9 - lse) == 0 in fp32 (the query's own slot is
// always selected and bounds lse from below), so the step only needs, for every slot i <= q_start:
//     valid(i)   : selected  <=>  #{valid j in (i, hi]} < K
//     otherwise  : selected  <=>  #{valid j in [0, hi]} + #{invalid j in (i, hi]} < K,      hi = min(q_start + 1, M - 1)
// (slot q_start + 1 is causally masked but competes for a place — kept, it is the reference's arithmetic).
// Three phases separated by block barriers: segment counts, (the caller's barrier), ranks inside the segment.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LSH_HD __host__ __device__ __forceinline__
#else
#define LSH_HD inline
#endif

namespace lsh {

constexpr int PREDICT_THREADS = 256;

struct PredictSelect {
  const int32_t *buckets;   // this unit's bucket memory (nh, M), AFTER the query's ids were written at q_start (EA:2069-2071)
  const int32_t *qb;        // the query's bucket id per round (nh)
  int M, nh, q_start, k_sel;
};

LSH_HD int predict_hi(const PredictSelect &p) { return p.q_start + 1 < p.M - 1 ? p.q_start + 1 : p.M - 1; }
LSH_HD int predict_seg(const PredictSelect &p) { return (predict_hi(p) + 1 + PREDICT_THREADS - 1) / PREDICT_THREADS; }

// Phase 1 (thread tid): validity flags of its segment of [0, hi] into flags[], counts into seg_valid / seg_invalid[tid].
LSH_HD void predict_select_count(const PredictSelect &p, int tid, uint8_t *flags, int *seg_valid, int *seg_invalid) {
  const int hi = predict_hi(p), seg = predict_seg(p);
  const int lo_i = tid * seg, hi_i = (tid + 1) * seg < hi + 1 ? (tid + 1) * seg : hi + 1;
  int cv = 0, ci = 0;
  for (int i = lo_i; i < hi_i; ++i) {
    bool v = i == p.q_start;
    for (int r = 0; r < p.nh && !v; ++r) v = p.buckets[static_cast<int64_t>(r) * p.M + i] == p.qb[r];
    flags[i] = v ? 1 : 0;
    cv += v ? 1 : 0;
    ci += v ? 0 : 1;
  }
  seg_valid[tid] = cv;
  seg_invalid[tid] = ci;
}

// Phase 2 (thread tid, after a barrier): flags[i] := 1 iff slot i <= q_start is one of the K attended slots.
LSH_HD void predict_select_rank(const PredictSelect &p, int tid, uint8_t *flags, const int *seg_valid, const int *seg_invalid) {
  const int hi = predict_hi(p), seg = predict_seg(p);
  const int lo_i = tid * seg, hi_i = (tid + 1) * seg < hi + 1 ? (tid + 1) * seg : hi + 1;
  int total_valid = 0, rv = 0, ri = 0;            // valid slots in [0, hi]; valid / invalid slots after this segment
  for (int t = 0; t < PREDICT_THREADS; ++t) {
    total_valid += seg_valid[t];
    if (t > tid) { rv += seg_valid[t]; ri += seg_invalid[t]; }
  }
  for (int i = hi_i - 1; i >= lo_i; --i) {
    const bool v = flags[i] != 0;
    const bool sel = v ? rv < p.k_sel : total_valid + ri < p.k_sel;
    if (v) ++rv; else ++ri;
    flags[i] = (sel && i <= p.q_start) ? 1 : 0;
  }
}

}  // namespace lsh
