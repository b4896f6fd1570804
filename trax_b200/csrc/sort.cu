// Stable bucket sort — replaces the two `sort_key_val` calls of EA:1946-1956.
//
// The reference sorts the int32 key  seqlen*bucket + position.  Bucket ids carry a per-round offset
// (EA:1913-1915), so hash round r owns sorted slots [r*L, (r+1)*L) and, inside a round, the order is
// "by bucket, then by position" = a STABLE sort of positions 0..L-1 by (bucket - r*n_buckets).  The
// permutation is unique (keys are unique as long as the int32 key does not wrap, which
// lsh_attn_check_dims enforces), so a stable counting sort reproduces jnp.argsort bit for bit.
//
// Each (unit, round) segment is an LSD radix sort with <= 11-bit digits (one pass when
// n_buckets <= 2048):   histogram per tile  ->  exclusive scan over (digit, tile)  ->  stable scatter
// (per-warp digit counters in shared memory + __match_any_sync ranks).  sticker[slot] = ticker of the
// element in that slot; undo_sort = inverse permutation (EA:1953).
#include "common.cuh"

namespace lsh {

constexpr int SORT_THREADS = 512;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_TILE = 4096;                       // positions per CTA
constexpr int SORT_PER_WARP = SORT_TILE / SORT_WARPS; // contiguous positions per warp
constexpr int SORT_MAX_BITS = 11;

struct SortParams {
  const int32_t *buckets; int64_t buckets_stride;
  const int32_t *perm_in;   // (BH, N) ticker values, or null = identity (first pass)
  int32_t *perm_out;        // (BH, N)
  int32_t *undo;            // (BH, N) or null; written on the last pass only
  int32_t *hist;            // (n_seg, n_digits, n_tiles)
  int L, nh, N, n_buckets, n_tiles, shift, n_digits, last_pass;
};

__device__ __forceinline__ int sort_digit(const SortParams &p, int u, int round, int idx, int &pos) {
  // idx: index inside the segment in the CURRENT order
  if (p.perm_in) pos = p.perm_in[static_cast<int64_t>(u) * p.N + round * p.L + idx] - round * p.L;
  else pos = idx;
  int b = p.buckets[static_cast<int64_t>(u) * p.buckets_stride + static_cast<int64_t>(round) * p.L + pos] -
          round * p.n_buckets;
  b = min(max(b, 0), p.n_buckets - 1);   // a corrupt state must not index outside shared memory
  return (b >> p.shift) & (p.n_digits - 1);
}

__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const SortParams p) {
  extern __shared__ int s_hist[];   // [n_digits]
  const int tile = blockIdx.x, seg = blockIdx.y;
  const int u = seg / p.nh, round = seg % p.nh;
  for (int i = threadIdx.x; i < p.n_digits; i += SORT_THREADS) s_hist[i] = 0;
  __syncthreads();
  const int base = tile * SORT_TILE;
  for (int i = threadIdx.x; i < SORT_TILE; i += SORT_THREADS) {
    int idx = base + i;
    if (idx < p.L) {
      int pos;
      int dgt = sort_digit(p, u, round, idx, pos);
      atomicAdd(&s_hist[dgt], 1);
    }
  }
  __syncthreads();
  int32_t *out = p.hist + static_cast<int64_t>(seg) * p.n_digits * p.n_tiles;
  for (int i = threadIdx.x; i < p.n_digits; i += SORT_THREADS)
    out[static_cast<int64_t>(i) * p.n_tiles + tile] = s_hist[i];
}

// Exclusive scan of one segment's (digit-major, tile-minor) histogram, in place.  The histogram is staged in shared
// memory with coalesced 16-byte accesses (33-word padding per 32 entries keeps the per-thread runs conflict-free);
// each thread then scans a contiguous run, a block scan links the runs.
__global__ void __launch_bounds__(1024) sort_scan_kernel(int32_t *hist, int entries) {
  extern __shared__ int s_h[];          // entries + entries/32 words
  __shared__ int s_warp[32];
  int32_t *h = hist + static_cast<int64_t>(blockIdx.x) * entries;
  const int per = (entries + 1023) / 1024;
  for (int i = threadIdx.x * 4; i < entries; i += 4096) {
    if ((entries & 3) == 0 && i + 3 < entries) {
      const int4 v = *reinterpret_cast<const int4 *>(h + i);
      s_h[i + (i >> 5)] = v.x; s_h[i + 1 + ((i + 1) >> 5)] = v.y;
      s_h[i + 2 + ((i + 2) >> 5)] = v.z; s_h[i + 3 + ((i + 3) >> 5)] = v.w;
    } else {
      for (int k = i; k < min(i + 4, entries); ++k) s_h[k + (k >> 5)] = h[k];
    }
  }
  __syncthreads();
  const int lo = min(threadIdx.x * per, entries), hi = min(lo + per, entries);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += s_h[i + (i >> 5)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    s_warp[lane] = wi - w;   // exclusive
  }
  __syncthreads();
  int run = s_warp[warp] + incl - sum;
  for (int i = lo; i < hi; ++i) {
    const int v = s_h[i + (i >> 5)];
    s_h[i + (i >> 5)] = run;
    run += v;
  }
  __syncthreads();
  for (int i = threadIdx.x * 4; i < entries; i += 4096) {
    if ((entries & 3) == 0 && i + 3 < entries) {
      int4 v;
      v.x = s_h[i + (i >> 5)]; v.y = s_h[i + 1 + ((i + 1) >> 5)];
      v.z = s_h[i + 2 + ((i + 2) >> 5)]; v.w = s_h[i + 3 + ((i + 3) >> 5)];
      *reinterpret_cast<int4 *>(h + i) = v;
    } else {
      for (int k = i; k < min(i + 4, entries); ++k) h[k] = s_h[k + (k >> 5)];
    }
  }
}

// Fallback for segments whose histogram does not fit in shared memory (e.g. 1M-token sequences): same scan,
// straight from global memory (each thread owns a contiguous run).
__global__ void __launch_bounds__(1024) sort_scan_global_kernel(int32_t *hist, int entries) {
  __shared__ int s_warp[32];
  int32_t *h = hist + static_cast<int64_t>(blockIdx.x) * entries;
  const int per = (entries + 1023) / 1024;
  const int lo = min(threadIdx.x * per, entries), hi = min(lo + per, entries);
  int sum = 0;
  for (int i = lo; i < hi; ++i) sum += h[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += v;
    }
    s_warp[lane] = wi - w;
  }
  __syncthreads();
  int run = s_warp[warp] + incl - sum;
  for (int i = lo; i < hi; ++i) {
    const int v = h[i];
    h[i] = run;
    run += v;
  }
}

__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const SortParams p) {
  extern __shared__ int s_cnt[];   // [SORT_WARPS][n_digits]
  const int tile = blockIdx.x, seg = blockIdx.y;
  const int u = seg / p.nh, round = seg % p.nh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < SORT_WARPS * p.n_digits; i += SORT_THREADS) s_cnt[i] = 0;
  __syncthreads();
  int *my_cnt = s_cnt + warp * p.n_digits;
  const int wbase = tile * SORT_TILE + warp * SORT_PER_WARP;
  // phase A: per-warp digit counts.  The digits (one dependent global load each) are fetched ONCE, all in flight
  // together, and kept in registers for phase C — its counter chain would otherwise serialise eight load latencies.
  constexpr int PER_LANE = SORT_PER_WARP / 32;
  int dg[PER_LANE], ps[PER_LANE];
#pragma unroll
  for (int k = 0; k < PER_LANE; ++k) {
    const int idx = wbase + k * 32 + lane;
    ps[k] = 0;
    dg[k] = (idx < p.L) ? sort_digit(p, u, round, idx, ps[k]) : -1;
  }
#pragma unroll
  for (int k = 0; k < PER_LANE; ++k)
    if (dg[k] >= 0) atomicAdd(&my_cnt[dg[k]], 1);
  __syncthreads();
  // phase B: counts -> starting offsets (tile base from the scanned histogram, then warps in order)
  const int32_t *hs = p.hist + static_cast<int64_t>(seg) * p.n_digits * p.n_tiles;
  for (int dgt = threadIdx.x; dgt < p.n_digits; dgt += SORT_THREADS) {
    int run = hs[static_cast<int64_t>(dgt) * p.n_tiles + tile];
#pragma unroll
    for (int w = 0; w < SORT_WARPS; ++w) {
      int c = s_cnt[w * p.n_digits + dgt];
      s_cnt[w * p.n_digits + dgt] = run;
      run += c;
    }
  }
  __syncthreads();
  // phase C: stable scatter, 32 elements of the warp's range at a time, in order
  const int64_t row = static_cast<int64_t>(u) * p.N + static_cast<int64_t>(round) * p.L;
#pragma unroll
  for (int k = 0; k < PER_LANE; ++k) {
    const bool valid = dg[k] >= 0;
    const int pos = ps[k];
    const unsigned dgt = static_cast<unsigned>(dg[k]);   // 0xffffffff for lanes past the end
    const unsigned peers = __match_any_sync(0xffffffffu, dgt);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int off = 0;
    if (valid) off = my_cnt[dgt];
    __syncwarp();
    if (valid && rank == 0) my_cnt[dgt] = off + __popc(peers);
    __syncwarp();
    if (valid) {
      const int dst = off + rank;                       // slot inside the round
      p.perm_out[row + dst] = round * p.L + pos;        // ticker value (EA:1946)
      if (p.last_pass && p.undo) p.undo[row + pos] = round * p.L + dst;
    }
  }
}

static int ceil_log2(int x) { int b = 0; while ((1 << b) < x) ++b; return b; }

struct SortPlan { int passes, bits, n_digits, n_tiles; size_t hist_bytes, perm_bytes; };

static SortPlan plan_sort(const LshAttnDims &d) {
  Derived dr = derive(d);
  SortPlan s;
  int total_bits = ceil_log2(dr.n_buckets);
  if (total_bits < 1) total_bits = 1;
  s.passes = (total_bits + SORT_MAX_BITS - 1) / SORT_MAX_BITS;
  s.bits = (total_bits + s.passes - 1) / s.passes;
  s.n_digits = 1 << s.bits;
  s.n_tiles = (d.L + SORT_TILE - 1) / SORT_TILE;
  s.hist_bytes = static_cast<size_t>(dr.BH) * d.nh * s.n_digits * s.n_tiles * sizeof(int32_t);
  s.hist_bytes = (s.hist_bytes + 255) / 256 * 256;
  s.perm_bytes = s.passes > 1 ? static_cast<size_t>(dr.BH) * dr.N * sizeof(int32_t) : 0;
  s.perm_bytes = (s.perm_bytes + 255) / 256 * 256;
  return s;
}

size_t sort_workspace_bytes(const LshAttnDims &d) {
  SortPlan s = plan_sort(d);
  return s.hist_bytes + s.perm_bytes;
}

int sort_run(const LshAttnDims &d, const int32_t *buckets, int64_t bstride, int32_t *sticker,
             int32_t *undo, void *ws, size_t ws_bytes, cudaStream_t stream) {
  Derived dr = derive(d);
  SortPlan s = plan_sort(d);
  if (ws_bytes < s.hist_bytes + s.perm_bytes)
    return set_error("lsh_sort: workspace too small (%zu < %zu)", ws_bytes, s.hist_bytes + s.perm_bytes);
  int32_t *hist = static_cast<int32_t *>(ws);
  int32_t *tmp = reinterpret_cast<int32_t *>(static_cast<char *>(ws) + s.hist_bytes);
  size_t smem_sc = static_cast<size_t>(SORT_WARPS) * s.n_digits * sizeof(int);
  LSH_OPT_IN_SMEM(sort_scatter_kernel);
  LSH_OPT_IN_SMEM(sort_scan_kernel);
  const int scan_entries = s.n_digits * s.n_tiles;
  const size_t scan_smem = static_cast<size_t>(scan_entries + scan_entries / 32 + 32) * sizeof(int);
  const bool scan_in_smem = scan_smem <= 200 * 1024;
  // ping-pong so that the last pass lands in `sticker`
  const int32_t *in = nullptr;
  for (int pass = 0; pass < s.passes; ++pass) {
    const bool last = pass == s.passes - 1;
    int32_t *out = ((s.passes - 1 - pass) % 2 == 0) ? sticker : tmp;
    SortParams p;
    p.buckets = buckets; p.buckets_stride = bstride; p.perm_in = in; p.perm_out = out;
    p.undo = undo; p.hist = hist; p.L = d.L; p.nh = d.nh; p.N = dr.N; p.n_buckets = dr.n_buckets;
    p.n_tiles = s.n_tiles; p.shift = pass * s.bits; p.n_digits = s.n_digits; p.last_pass = last;
    dim3 grid(s.n_tiles, dr.BH * d.nh);
    sort_hist_kernel<<<grid, SORT_THREADS, s.n_digits * sizeof(int), stream>>>(p);
    LSH_CHECK_LAUNCH("sort_hist_kernel");
    if (scan_in_smem) sort_scan_kernel<<<dr.BH * d.nh, 1024, scan_smem, stream>>>(hist, s.n_digits * s.n_tiles);
    else sort_scan_global_kernel<<<dr.BH * d.nh, 1024, 0, stream>>>(hist, s.n_digits * s.n_tiles);
    LSH_CHECK_LAUNCH("sort_scan_kernel");
    sort_scatter_kernel<<<grid, SORT_THREADS, smem_sc, stream>>>(p);
    LSH_CHECK_LAUNCH("sort_scatter_kernel");
    in = out;
  }
  return 0;
}

// Re-orders every 128-slot chunk of `sticker` by token position (ascending; ties — only possible when a chunk straddles
// two hash rounds — keep slot order).  Attention within a chunk window is a sum over keys and its rows are scattered
// back by ticker, so the order inside a chunk is free: with position-sorted tiles the causal mask of EA:150-152 becomes
// an interval of column indices per row, whole 32-column blocks are either fully visible or skipped, and no per-key
// position has to be loaded in the softmax loop.
//
// The same launch emits, per row (in the new order), where the row's position falls inside the two NEIGHBOUR chunks of
// its unit (cyclic, EA:137-141) — the interval bounds the attention kernels would otherwise find by an 8-step binary
// search per row and pass:
//   bounds = cnt_prev | eq_prev << 8 | cnt_next << 16 | eq_next << 24
//   cnt_prev / cnt_next = number of tokens of chunk c-1 / c+1 whose position is BELOW the row's; eq_* = that chunk holds
//   the row's own position (its copy from the neighbouring hash round).
// One warp per chunk (warp-level bitonic network on (position << 7 | slot) keys); a CTA covers 8 consecutive chunks of
// one unit and sorts the two halo chunks redundantly (10 warps).
constexpr int POSSORT_WARPS = 10;
__global__ void __launch_bounds__(32 * POSSORT_WARPS) chunk_possort_kernel(const int32_t *__restrict__ sticker,
                                                                          int32_t *__restrict__ sticker2,
                                                                          int32_t *__restrict__ bounds, int L, int n_chunks,
                                                                          int blocks_per_unit) {
  __shared__ int stk[POSSORT_WARPS][128];
  __shared__ int spos[POSSORT_WARPS][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x / blocks_per_unit, c0 = (blockIdx.x - u * blocks_per_unit) * 8;
  const int c_raw = c0 - 1 + warp;
  const int c = ((c_raw % n_chunks) + n_chunks) % n_chunks;                        // cyclic inside the unit
  const bool owner = warp >= 1 && warp <= 8 && c_raw < n_chunks;                  // this warp's chunk is written by this CTA
  if (bounds == nullptr && !owner) return;                                        // halo warps only serve the bounds
  const int64_t base = (static_cast<int64_t>(u) * n_chunks + c) * 128;
  // element e = 4 * lane + i; key = (position << 7) | slot: unique, so the bitonic network needs no tie rule
  const int4 t4 = __ldg(reinterpret_cast<const int4 *>(sticker + base) + lane);
  *reinterpret_cast<int4 *>(&stk[warp][4 * lane]) = t4;
  uint32_t key[4] = {static_cast<uint32_t>(t4.x % L) << 7 | (4 * lane + 0), static_cast<uint32_t>(t4.y % L) << 7 | (4 * lane + 1),
                     static_cast<uint32_t>(t4.z % L) << 7 | (4 * lane + 2), static_cast<uint32_t>(t4.w % L) << 7 | (4 * lane + 3)};
#pragma unroll
  for (int k = 2; k <= 128; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 4) {                                   // partner element lives in lane ^ (j / 4), same i
        const bool upper = (lane & (j >> 2)) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = 4 * lane + i;
          const bool desc = (e & k) != 0;
          const uint32_t other = __shfl_xor_sync(0xffffffffu, key[i], j >> 2);
          const bool take_max = upper != desc;      // ascending block: lower index keeps the min
          key[i] = take_max ? max(key[i], other) : min(key[i], other);
        }
      } else {                                        // both elements in this thread: i and i ^ j
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int i2 = i ^ j;
          if (i2 > i) {
            const int e = 4 * lane + i;
            const bool desc = (e & k) != 0;
            const uint32_t lo = min(key[i], key[i2]), hi = max(key[i], key[i2]);
            key[i] = desc ? hi : lo;
            key[i2] = desc ? lo : hi;
          }
        }
      }
    }
  }
  __syncwarp();
  if (owner) {
    int4 o4;
    o4.x = stk[warp][key[0] & 127]; o4.y = stk[warp][key[1] & 127]; o4.z = stk[warp][key[2] & 127]; o4.w = stk[warp][key[3] & 127];
    *(reinterpret_cast<int4 *>(sticker2 + base) + lane) = o4;
  }
  if (bounds == nullptr) return;
  *reinterpret_cast<int4 *>(&spos[warp][4 * lane]) = make_int4(static_cast<int>(key[0] >> 7), static_cast<int>(key[1] >> 7),
                                                               static_cast<int>(key[2] >> 7), static_cast<int>(key[3] >> 7));
  __syncthreads();
  if (!owner) return;
  // neighbours: warp - 1 holds chunk c - 1, warp + 1 chunk c + 1 (both cyclic; for a one-chunk unit all three coincide)
  const int *prev = spos[warp - 1], *next = spos[warp + 1];
  int4 b4;
  int *bo = reinterpret_cast<int *>(&b4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = static_cast<int>(key[i] >> 7);
    int lp = 0, ln = 0;                                 // lower bounds (number of entries below pos)
#pragma unroll
    for (int s = 64; s >= 1; s >>= 1) {
      if (prev[lp + s - 1] < pos) lp += s;
      if (next[ln + s - 1] < pos) ln += s;
    }
    if (lp == 127 && prev[127] < pos) lp = 128;
    if (ln == 127 && next[127] < pos) ln = 128;
    const int ep = (lp < 128 && prev[lp] == pos) ? 1 : 0, en = (ln < 128 && next[ln] == pos) ? 1 : 0;
    bo[i] = lp | (ep << 8) | (ln << 16) | (en << 24);
  }
  *(reinterpret_cast<int4 *>(bounds + base) + lane) = b4;
}

int chunk_possort_run(const LshAttnDims &d, const int32_t *sticker, int32_t *sticker2, int32_t *bounds, cudaStream_t stream) {
  Derived dr = derive(d);
  if (d.C != 128) return set_error("chunk_possort: chunk_len must be 128");
  const int bpu = (dr.n_chunks + 7) / 8;
  chunk_possort_kernel<<<static_cast<unsigned>(dr.BH * bpu), 32 * POSSORT_WARPS, 0, stream>>>(sticker, sticker2, bounds, d.L,
                                                                                             dr.n_chunks, bpu);
  LSH_CHECK_LAUNCH("chunk_possort_kernel");
  return 0;
}

}  // namespace lsh
