// Chunked look-back attention, backward, on tcgen05 + TMEM.  Same math as attend_bwd.cu (SURVEY App. B with
// the multi-round combine folded in); specialised for chunk_len 128, n_chunks_before = 1, n_chunks_after = 0,
// causal, no padding mask, dq = dv = 64 (every long-sequence config).
//
// KEY-centric persistent walk: one CTA per SM owns a contiguous range of key chunks.  For key chunk t it visits
// the two query chunks that see it (t itself and t+1), 64 queries ("half") at a time, in the transposed
// orientation so that the softmax threads own KEY rows (TMEM lane = key row):
//     S^T  = K_t Q^T           dP^T = V_t dO^T                       (tcgen05.mma SS, M128 N64 K64)
//     P^T  = [k_j < q_i] exp2(S^T ksc_j - lse2_i) ;  dS^T = P^T (dP^T - D_i)     (softmax warpgroups, in place)
//     dV  += P^T dO   ;  dK^ += dS^T Q                                (tcgen05.mma TS, A from TMEM)
//     dQ  += (dS ksc) K_t      (A = dS staged in shared memory, MN-major)         (tcgen05.mma SS)
// dK^ / dV of chunk t are complete after its two query chunks (no atomics, no partial buffers); dQ of chunk t
// collects keys t-1 (previous iteration) and t (this iteration) in one of two TMEM slots.  Because queries
// and keys are the same tokens (shared-QK), thread j ends up holding dq_query, dq_key and dv of ONE token and
// writes single rows.  Range / unit boundaries replay one "pre" item (keys t-1 x queries t, dQ only).
//
// TMEM (512 columns): region h (h = query half): S^T [128h,+64) -> P^T bf16 [128h,+32); dP^T [128h+64,+64) ->
// dS^T bf16 [128h+64,+32); dK^ [256,320); dV [320,384); dQ slots [384,448), [448,512).
// Warps: 0-3 softmax warpgroup 0 (half 0), 4-7 warpgroup 1 (half 1), 8-11 epilogue warpgroup (drains dK^/dV first — four TMEM
// loads, then kv_free lets the next key chunk accumulate — and dQ afterwards, guarded by dq_free), 12-13 producers
// (position-sorted sticker2, copy-signalled tiles), 14 MMA issuer (pre-issues the next item's S^T / dP^T as soon as a
// non-blocking probe finds its tiles landed).
// Tiles are ordered by position, so per warp a 32-query block is skipped (zeros), evaluated without the position compare,
// or — around the boundary — evaluated with it; -lse2 and -D are stored negated so the math runs on packed fp32 pairs.
#include "attend_bwd_params.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

// k-step loops of the MMA issuer: fully unrolled by default (immediate descriptor offsets); -DLSH_BWD_ROLL_ISSUE rolls
// them (A/B hook for the issuer's instruction footprint).
#ifdef LSH_BWD_ROLL_ISSUE
#define BT_ISSUE_UNROLL _Pragma("unroll 1")
#else
#define BT_ISSUE_UNROLL _Pragma("unroll")
#endif

namespace lsh {

constexpr int BT_C = 128;
constexpr int BT_NST = 3;                         // tile ring depth
constexpr int BT_THREADS = 480;
constexpr int BT_TILE_BYTES = 3 * BT_C * 128;     // K(=Q) rows | V rows | dO rows
constexpr int BT_DS_BYTES = 2 * BT_C * 128;       // one dS staging buffer: two 64-query blocks of [128 keys][128 B]
constexpr uint32_t BT_IDESC_ST = make_idesc_bf16(128, 64, 0, 0);   // S^T, dP^T
constexpr uint32_t BT_IDESC_KV = make_idesc_bf16(128, 64, 0, 1);   // dV, dK^ (B MN-major)
constexpr uint32_t BT_IDESC_DQ = make_idesc_bf16(128, 64, 1, 1);   // dQ (A and B MN-major)

// trace slots per item (32): 0 MMA: pds_full[0] seen, 1 MMA: pds_full[1] seen, 2 MMA: item done (commits issued),
// 3..10 softmax warp w: pass done, 11..18 softmax warp w: st_full seen, 19/21 MMA: dV/dK of half 0/1 issued, 20/22 MMA: next
// item's S^T half 0/1 issued, 23 MMA: dQ issued, 24 epilogue done, 25 epilogue: kv_full seen, 26 producer: tile issued,
// 27 softmax warp 0: tiles seen
// (compiled in only with -DLSH_TRACE: the stamps cost instruction-cache space in every role)
#ifdef LSH_TRACE
#define BT_TRACE(n, slot) do { if (p.trace && blockIdx.x == 0 && (n) < 120) p.trace[(n) * 32 + (slot)] = clock64(); } while (0)
#else
#define BT_TRACE(n, slot) do { } while (0)
#endif

struct __align__(16) BtTileMeta {
  float kinfo[BT_C];    // pos + 1 as fp32 (key side of the causal compare)
  float qcmp[BT_C];     // pos + 1, or pos + 1.5 for rows whose only visible key is their own "-1e5" class
  float lse2[BT_C];     // -(log2(e) * lse_tot (+ the -1e5 class shift for those rows))
  float dvec[BT_C];     // -D_i = -(do_i . o_i)
  float kscl[BT_C];     // log2(e) / (sqrt(mean(q^2)+eps) * sqrt(dq))
  int tk[BT_C];         // ticker
};

struct __align__(16) BtShared {
  BtTileMeta meta[BT_NST];
  uint64_t full[BT_NST], empty[BT_NST];
  uint64_t st_full[2], pds_full[2], dsm_free[2], dq_full[2];
  uint64_t kv_full, kv_free, dq_free[2];
  uint32_t tmem_base;
};

// One (key tile, query tile) product group.
struct BtItem {
  int seq_k, seq_q;       // tile sequence numbers
  int u;                  // unit of the key chunk
  int n;                  // global item index (barrier phases)
  int rit;                // index of the real iteration this item belongs to (or the upcoming one for a pre item)
  bool real, first, iter_end, last_seg;   // first = first item of its iteration
  bool do_kv, kv_first, do_dq, dq_fresh;
  int dq_slot;
};

// Enumerates iterations / items / tiles of one CTA's range of key chunks [g0, g1).
struct BtWalk {
  int g, g_end, nc, seq, rit, n, u, c;
  bool need_pre, second;      // second = item B of a real iteration comes next
  __device__ BtWalk(int g0, int g1, int nchunks)
      : g(g0), g_end(g1), nc(nchunks), seq(0), rit(0), n(0), need_pre(true), second(false) {
    u = g0 / nchunks; c = g0 - u * nchunks;
  }
  __device__ bool valid() const { return g < g_end; }
  __device__ bool seg_last() const { return g == g_end - 1 || c == nc - 1; }
  __device__ BtItem item() const {
    BtItem it;
    it.u = u; it.n = n; it.rit = rit; it.seq_k = seq;
    if (need_pre) {
      it.seq_q = seq + 1; it.real = false; it.first = true; it.iter_end = true; it.last_seg = false;
      it.do_kv = false; it.kv_first = false; it.do_dq = true; it.dq_fresh = true; it.dq_slot = rit & 1;
    } else if (!second) {
      it.seq_q = seq; it.real = true; it.first = true; it.iter_end = false; it.last_seg = seg_last();
      it.do_kv = true; it.kv_first = true; it.do_dq = true; it.dq_fresh = false; it.dq_slot = rit & 1;
    } else {
      it.seq_q = seq + 1; it.real = true; it.first = false; it.iter_end = true; it.last_seg = seg_last();
      it.do_kv = true; it.kv_first = false; it.do_dq = !it.last_seg; it.dq_fresh = true; it.dq_slot = (rit + 1) & 1;
    }
    return it;
  }
  __device__ void next() {
    ++n;
    if (need_pre) { need_pre = false; seq += 1; return; }
    if (!second) { second = true; return; }
    second = false;
    const bool last = seg_last();
    seq += last ? 2 : 1;
    need_pre = last;
    ++rit; ++g; ++c;
    if (c == nc) { c = 0; ++u; }
  }
  // chunk ids (within the unit) of the key tile and of the next tile of the current iteration
  __device__ void chunks(int &c_key, int &c_next) const {
    if (need_pre) { c_key = (c == 0) ? nc - 1 : c - 1; c_next = c; } else { c_key = c; c_next = (c + 1 == nc) ? 0 : c + 1; }
  }
};

__device__ __forceinline__ uint32_t bt_slot(int seq) { return static_cast<uint32_t>(seq % BT_NST); }
__device__ __forceinline__ uint32_t bt_phase(int seq) { return static_cast<uint32_t>((seq / BT_NST) & 1); }

// DROPOUT: slot-ordered tiles + keep bits (see AttendBwdTcParams); its own instantiation, so that the plain kernel's code —
// register allocation, instruction-cache footprint — is exactly the one without it (measured: +2.5 % otherwise).
template <bool DROPOUT>
__global__ void __launch_bounds__(BT_THREADS, 1) attend_bwd_tc_kernel(const AttendBwdTcParams p, int total_chunks) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *tiles = smem;                                   // [BT_NST][K | V | dO]
  uint8_t *dsbuf = smem + BT_NST * BT_TILE_BYTES;          // [2][BT_DS_BYTES]
  __shared__ BtShared sh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0 = static_cast<int>(static_cast<int64_t>(total_chunks) * blockIdx.x / gridDim.x);
  const int g1 = static_cast<int>(static_cast<int64_t>(total_chunks) * (blockIdx.x + 1) / gridDim.x);

  if (warp == 14) tmem_alloc(&sh.tmem_base, 512);
  if (tid == 0) {
    // full: per producer thread one arrival by its copies (noinc) and one ordinary arrival releasing its metadata stores
    for (int i = 0; i < BT_NST; ++i) { mbar_init(&sh.full[i], 128); mbar_init(&sh.empty[i], 385); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh.st_full[i], 1); mbar_init(&sh.pds_full[i], 128);
      mbar_init(&sh.dsm_free[i], 1); mbar_init(&sh.dq_full[i], 1);
    }
    mbar_init(&sh.kv_full, 1); mbar_init(&sh.kv_free, 128);
    mbar_init(&sh.dq_free[0], 128); mbar_init(&sh.dq_free[1], 128);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  const uint32_t tiles_u32 = smem_u32(tiles), ds_u32 = smem_u32(dsbuf);

  if (warp == 12 || warp == 13) {
    // ================================ producers ========================================================
    const int pw = warp - 12;                              // rows [64*pw, 64*pw + 64) of every tile
    auto load_tile = [&](int seq, int u, int cc) {
      const uint32_t slot = bt_slot(seq);
      const int b = u / p.H, h = u - b * p.H;
      const int32_t *stk = p.sticker2 + static_cast<int64_t>(u) * p.N + cc * BT_C + 64 * pw;   // chunk rows ascending in position
      const int tka = __ldg(stk + lane), tkb = __ldg(stk + 32 + lane);
      const int pa = tka % p.L, pb = tkb % p.L;
      mbar_wait<128>(&sh.empty[slot], bt_phase(seq) ^ 1);
      BtTileMeta &mt = sh.meta[slot];
      const int ra = 64 * pw + lane, rb = ra + 32;
      mt.kinfo[ra] = static_cast<float>(pa + 1); mt.kinfo[rb] = static_cast<float>(pb + 1);
      mt.tk[ra] = tka; mt.tk[rb] = tkb;
      const int64_t oa = static_cast<int64_t>(u) * p.L + pa, ob = static_cast<int64_t>(u) * p.L + pb;
      cp_async4(smem_u32(&mt.kscl[ra]), p.qscale + oa); cp_async4(smem_u32(&mt.kscl[rb]), p.qscale + ob);
      cp_async4(smem_u32(&mt.lse2[ra]), p.lse2 + oa);   cp_async4(smem_u32(&mt.lse2[rb]), p.lse2 + ob);
      cp_async4(smem_u32(&mt.dvec[ra]), p.dvec + oa);   cp_async4(smem_u32(&mt.dvec[rb]), p.dvec + ob);
      cp_async4(smem_u32(&mt.qcmp[ra]), p.qcmp + oa);   cp_async4(smem_u32(&mt.qcmp[rb]), p.qcmp + ob);
      const uint32_t kt = tiles_u32 + slot * BT_TILE_BYTES, vt = kt + BT_C * 128, dt = vt + BT_C * 128;
      {   // q|v rows: 16 lanes per 256-byte row, 2 rows per instruction.  Destination of row R = 64 pw + 2 i + hi, piece c:
          // tile + (R << 7) + (((c ^ R) & 7) << 4) = tile + (lane_const ^ imm(i)) — all bit fields are disjoint.
        const int ch = lane & 15, hi = lane >> 4;
        const char *base = reinterpret_cast<const char *>(p.qv + (static_cast<int64_t>(b) * p.L * p.H + h) * 128 + ch * 8);
        const uint32_t rbytes = static_cast<uint32_t>(p.H) * 256u;
        const uint32_t tile = ch < 8 ? kt : vt;
        const uint32_t lane_const = (static_cast<uint32_t>(64 * pw + hi) << 7) | (static_cast<uint32_t>((ch ^ hi) & 7) << 4);
#pragma unroll 1
        for (int io = 0; io < 8; ++io) {
          const int psel = io < 4 ? pa : pb;
          const uint32_t dst = tile + (lane_const ^ (static_cast<uint32_t>(io) << 10));
#pragma unroll
          for (int ii = 0; ii < 4; ++ii) {
            const int pr = __shfl_sync(0xffffffffu, psel, (8 * io + 2 * ii + hi) & 31);
            cp_async16(dst ^ ((static_cast<uint32_t>(ii) << 8) ^ (static_cast<uint32_t>(ii) << 5)), base + static_cast<uint64_t>(static_cast<uint32_t>(pr)) * rbytes);
          }
        }
      }
      {   // do rows: 8 lanes per 128-byte row, 4 rows per instruction: R = 64 pw + 4 i + hi (hi = 0..3)
        const int ch = lane & 7, hi = lane >> 3;
        const char *base = reinterpret_cast<const char *>(p.do_comb + (static_cast<int64_t>(b) * p.L * p.H + h) * 64 + ch * 8);
        const uint32_t rbytes = static_cast<uint32_t>(p.H) * 128u;
        const uint32_t lane_const = (static_cast<uint32_t>(64 * pw + hi) << 7) | (static_cast<uint32_t>((ch ^ hi) & 7) << 4);
#pragma unroll 1
        for (int io = 0; io < 8; ++io) {                  // i = 2 io + ii: rows 8 io + 4 ii + hi
          const int psel = io < 4 ? pa : pb;
          const uint32_t dst = dt + (lane_const ^ (static_cast<uint32_t>(io) << 10));
#pragma unroll
          for (int ii = 0; ii < 2; ++ii) {
            const int pr = __shfl_sync(0xffffffffu, psel, (8 * io + 4 * ii + hi) & 31);
            cp_async16(dst ^ ((static_cast<uint32_t>(ii) << 9) ^ (static_cast<uint32_t>(ii) << 6)), base + static_cast<uint64_t>(static_cast<uint32_t>(pr)) * rbytes);
          }
        }
      }
      // Completion is signalled by the copies themselves (cp.async.mbarrier.arrive.noinc): the producer never waits for
      // its own loads, and the tile becomes visible the moment it has landed.
      cp_async_mbar_arrive_noinc(&sh.full[slot]);
      mbar_arrive(&sh.full[slot]);                         // (release of the plain kinfo / tk stores above)
    };
    for (BtWalk w(g0, g1, p.n_chunks); w.valid();) {
      const BtItem it = w.item();
      if (it.first) {               // one visit per iteration: a pre iteration brings two tiles, a real one brings one
        int c_key, c_next;
        w.chunks(c_key, c_next);
        if (!it.real) load_tile(it.seq_k, it.u, c_key);
        load_tile(it.seq_k + 1, it.u, c_next);
        if (tid == 12 * 32) BT_TRACE(it.n, 26);
      }
      w.next();
    }
  } else if (warp == 14) {
    // ================================ MMA issuer =======================================================
    // The whole warp runs this code converged (all values warp-uniform); one elected lane issues the
    // tcgen05.mma / tcgen05.commit instructions.  Descriptors: constant hi word, lo = base + (offset >> 4).
    constexpr uint32_t HI = desc_hi(1024);
    auto tile_addr = [&](int seq) { return tiles_u32 + bt_slot(seq) * BT_TILE_BYTES; };
    // S^T and dP^T of one half of an item
    auto issue_st = [&](const BtItem &it, int h) {
      const uint32_t kt = tile_addr(it.seq_k), qt = tile_addr(it.seq_q);
      const uint32_t r = tmem + 128 * h;
      const uint32_t ka = desc_lo(kt, 16), qb = desc_lo(qt + h * 8192, 16);
      const uint32_t va = desc_lo(kt + BT_C * 128, 16), db = desc_lo(qt + 2 * BT_C * 128 + h * 8192, 16);
      if (elect_one()) {
BT_ISSUE_UNROLL
        for (int ks = 0; ks < 4; ++ks) umma_ss2(r, ka + ks * 2, HI, qb + ks * 2, HI, BT_IDESC_ST, ks > 0);
BT_ISSUE_UNROLL
        for (int ks = 0; ks < 4; ++ks) umma_ss2(r + 64, va + ks * 2, HI, db + ks * 2, HI, BT_IDESC_ST, ks > 0);
        umma_commit(&sh.st_full[h]);
      }
      __syncwarp();
    };
    auto wait_tiles = [&](const BtItem &it) {
      mbar_wait(&sh.full[bt_slot(it.seq_k)], bt_phase(it.seq_k));
      mbar_wait(&sh.full[bt_slot(it.seq_q)], bt_phase(it.seq_q));
      tc_fence_after();   // (tiles are signalled by their cp.async copies; same protocol as CUTLASS's sm100 cp.async mainloop)
    };
    int n_epi[2] = {0, 0};                                   // epilogues started per dQ slot (kv_full commits)
    BtWalk w(g0, g1, p.n_chunks);
    BtItem cur = w.item();
    wait_tiles(cur);
    issue_st(cur, 0);
    issue_st(cur, 1);
    while (true) {
      w.next();
      const bool have_next = w.valid();
      BtItem nxt = cur;
      if (have_next) nxt = w.item();
      const uint32_t qt = tile_addr(cur.seq_q), kt = tile_addr(cur.seq_k);
      // Pre-issue the next item's S^T / dP^T inside this item as soon as its tiles have landed — probed (never waited
      // for: at a segment boundary the next tiles reuse the slot of a tile this item still holds) before each half and
      // once more after the dQ MMAs, because at an iteration boundary the new key tile typically lands late in the item.
      bool pre_ok = false;
      int st_issued = 0;                                     // halves of the next item already issued
      auto probe = [&]() {
        if (have_next && !pre_ok) {
          bool ok = mbar_test(&sh.full[bt_slot(nxt.seq_k)], bt_phase(nxt.seq_k)) &&
                    mbar_test(&sh.full[bt_slot(nxt.seq_q)], bt_phase(nxt.seq_q));
          ok = __all_sync(0xffffffffu, ok);
          if (ok) { tc_fence_after(); pre_ok = true; }
        }
      };
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t r = tmem + 128 * h;
        const uint32_t dob = desc_lo(qt + 2 * BT_C * 128 + h * 8192, 1024), qb = desc_lo(qt + h * 8192, 1024);
        mbar_wait(&sh.pds_full[h], cur.n & 1);
        tc_fence_after();
        if (lane == 0) BT_TRACE(cur.n, h);
        if (cur.do_kv) {
          if (cur.kv_first && h == 0) mbar_wait(&sh.kv_free, (cur.rit & 1) ^ 1);   // epilogue warpgroup has drained dK^/dV/dQ
          const uint32_t fresh = (cur.kv_first && h == 0) ? 1u : 0u;
          if (elect_one()) {
BT_ISSUE_UNROLL
            for (int kk = 0; kk < 4; ++kk) umma_ts2(tmem + 320, r + kk * 8, dob + kk * 128, HI, BT_IDESC_KV, !(fresh && kk == 0));
BT_ISSUE_UNROLL
            for (int kk = 0; kk < 4; ++kk) umma_ts2(tmem + 256, r + 64 + kk * 8, qb + kk * 128, HI, BT_IDESC_KV, !(fresh && kk == 0));
          }
          __syncwarp();
        }
        if (lane == 0) BT_TRACE(cur.n, 19 + 2 * h);
        // region h is free once the MMAs above have read it (in-order pipe)
        probe();
        if (pre_ok) {
          for (; st_issued <= h; ++st_issued) issue_st(nxt, st_issued);
          if (lane == 0) BT_TRACE(cur.n, 20 + 2 * h);
        }
      }
      {
        const uint32_t a0 = desc_lo(ds_u32 + (cur.n & 1) * BT_DS_BYTES, 16384), kb = desc_lo(kt, 1024);
        const uint32_t dq_t = tmem + 384 + 64 * cur.dq_slot;
        // a fresh dQ accumulation re-uses the slot the epilogue of two iterations back drained (it signals dq_free as
        // soon as its dQ loads are done; dK^/dV are handed back separately and earlier through kv_free)
        if (cur.do_dq && cur.dq_fresh && n_epi[cur.dq_slot] > 0) mbar_wait(&sh.dq_free[cur.dq_slot], (n_epi[cur.dq_slot] - 1) & 1);
        if (cur.do_kv && cur.iter_end) ++n_epi[cur.rit & 1];
        if (elect_one()) {
          if (cur.do_kv && cur.iter_end) umma_commit(&sh.kv_full);
          if (cur.do_dq) {
BT_ISSUE_UNROLL
            for (int kk = 0; kk < 8; ++kk) umma_ss2(dq_t, a0 + kk * 128, HI, kb + kk * 128, HI, BT_IDESC_DQ, !(cur.dq_fresh && kk == 0));
          }
          umma_commit(&sh.dsm_free[cur.n & 1]);
          if (cur.real && !cur.iter_end) umma_commit(&sh.dq_full[cur.rit & 1]);     // dQ of this key chunk is final after item A
          if (cur.iter_end) {
            umma_commit(&sh.empty[bt_slot(cur.seq_k)]);
            if (cur.real && cur.last_seg) umma_commit(&sh.empty[bt_slot(cur.seq_k + 1)]);
          }
        }
        __syncwarp();
      }
      if (lane == 0) BT_TRACE(cur.n, 23);
      probe();
      if (pre_ok) {
        for (; st_issued < 2; ++st_issued) issue_st(nxt, st_issued);
      }
      if (lane == 0) BT_TRACE(cur.n, 2);
      if (!have_next) break;
      if (st_issued < 2) {
        wait_tiles(nxt);
        for (; st_issued < 2; ++st_issued) issue_st(nxt, st_issued);
      }
      cur = nxt;
    }
  } else if (warp < 8) {
    // ================================ softmax warpgroups ===============================================
    const int h = warp >> 2;                               // query half handled by this warpgroup
    const int row = (warp & 3) * 32 + lane;                // key row == TMEM lane (lane quarter = warp id % 4)
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t r_st = t_lane + 128 * h;
    float ksc_j = 0.f, ki_j = 0.f;
    // my 128-byte row of the dS staging tile (shared-space address: plain STS, no generic-address stores)
    const uint32_t ds_row = ds_u32 + h * (BT_C * 128) + row * 128, r7 = static_cast<uint32_t>(row & 7);
    for (BtWalk w(g0, g1, p.n_chunks); w.valid(); w.next()) {
      const BtItem it = w.item();
      const uint32_t slk = bt_slot(it.seq_k), slq = bt_slot(it.seq_q);
      if (it.first) {
        mbar_wait(&sh.full[slk], bt_phase(it.seq_k));
        ksc_j = sh.meta[slk].kscl[row];
        ki_j = sh.meta[slk].kinfo[row];
      }
      mbar_wait(&sh.full[slq], bt_phase(it.seq_q));
      if (tid == 0) BT_TRACE(it.n, 27);
      const BtTileMeta &mq = sh.meta[slq];
      const float kst_j = ksc_j * kLn2;                    // true key scale 1/(r*sqrt(dq)) for the dQ operand
      const uint64_t ksc2 = pk2(ksc_j, ksc_j), kst2 = pk2(kst_j, kst_j);
      const uint32_t ds_dst = ds_row + (it.n & 1) * BT_DS_BYTES;
      // Tiles are ordered by position (chunk_possort_kernel), so for key j the queries of the tile split into three index
      // ranges: [0, lo) position below the key's — never visible (EA:150-152); [lo, hi) the same position — the key's own
      // token or its copy from the neighbouring hash round, visible only to self-only rows (qcmp rule); [hi, 128) visible.
      // Per warp, a 32-query block is skipped (zeros), evaluated without any mask (no per-query position loads), or
      // — only around the boundary — evaluated with the per-element compare.
      int lo_j, hi_j;
      if (it.seq_q == it.seq_k) {
        lo_j = row; hi_j = row + 1;
      } else {
        int blo = 0, bhi = 128;
#pragma unroll
        for (int sidx = 0; sidx < 8; ++sidx) {
          const int mid = (blo + bhi) >> 1;
          const float v = mq.kinfo[mid & 127];
          const bool go = blo < bhi;
          if (go && v < ki_j) blo = mid + 1;
          else if (go) bhi = mid;
        }
        lo_j = blo;
        hi_j = blo + ((blo < 128 && mq.kinfo[blo & 127] == ki_j) ? 1 : 0);
      }
      if (DROPOUT) { lo_j = 0; hi_j = BT_C; }                // tiles in slot order (dropout): every block with the compare
      const int min_lo = __reduce_min_sync(0xffffffffu, lo_j), max_hi = __reduce_max_sync(0xffffffffu, hi_j);
      // dropout: my key's row of the transposed keep matrix — bit i = query slot i keeps this key (window column = my slot in
      // the look-back part, C + my slot in the own part); all ones and scale 1 without dropout
      uint32_t kw0 = 0xffffffffu, kw1 = 0xffffffffu;
      float kscale = 1.f;
      if (DROPOUT) {
        const uint4 kw = __ldg(reinterpret_cast<const uint4 *>(p.keep_bits_t) + ((it.seq_q == it.seq_k ? BT_C : 0) + row));
        kw0 = h ? kw.z : kw.x; kw1 = h ? kw.w : kw.y;
        kscale = __ldg(p.keep_scale);
      }
      mbar_wait(&sh.st_full[h], it.n & 1);
      mbar_wait(&sh.dsm_free[it.n & 1], ((it.n >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) BT_TRACE(it.n, 11 + warp);
#pragma unroll 1
      for (int cc = 0; cc < 64; cc += 32) {
        const int c0 = 64 * h + cc;
        const uint32_t kwb = cc ? kw1 : kw0;                  // keep bits of this block's 32 queries
#ifdef LSH_TRACE_WARP
        if (cc == 32 && warp == LSH_TRACE_WARP && lane == 0) BT_TRACE(it.n, 31);
#endif
        uint32_t s[32], pk_p[16], pk_ds[16];
        if (c0 + 32 <= min_lo) {
          // no key of this warp sees any query of the block
#pragma unroll
          for (int i = 0; i < 16; ++i) { pk_p[i] = 0u; s[i] = 0u; }
          tmem_st16(r_st + (cc >> 1), pk_p);
          tmem_st16(r_st + 64 + (cc >> 1), pk_p);
        } else {
          uint32_t dp[32];
          tmem_ld32(r_st + cc, s);
          tmem_ld32(r_st + 64 + cc, dp);
          tmem_ld_wait_dep(s);
          tmem_ld_wait_dep(dp);
          if (c0 >= max_hi) {
            // every key of this warp sees every query of the block
#pragma unroll
            for (int c4 = 0; c4 < 32; c4 += 4) {
              const ulonglong2 ls = *reinterpret_cast<const ulonglong2 *>(&mq.lse2[c0 + c4]);   // -lse2 pairs
              const ulonglong2 dv = *reinterpret_cast<const ulonglong2 *>(&mq.dvec[c0 + c4]);   // -D pairs
              const uint64_t t01 = ffma2(pk2u(s[c4 + 0], s[c4 + 1]), ksc2, ls.x), t23 = ffma2(pk2u(s[c4 + 2], s[c4 + 3]), ksc2, ls.y);
              const uint64_t p01 = pk2(fast_exp2(lo32(t01)), fast_exp2(hi32(t01))), p23 = pk2(fast_exp2(lo32(t23)), fast_exp2(hi32(t23)));
              const uint64_t d01 = fmul2(p01, fadd2(pk2u(dp[c4 + 0], dp[c4 + 1]), dv.x));
              const uint64_t d23 = fmul2(p23, fadd2(pk2u(dp[c4 + 2], dp[c4 + 3]), dv.y));
              const uint64_t g01 = fmul2(d01, kst2), g23 = fmul2(d23, kst2);
              pk_p[c4 >> 1] = pack_bf16(lo32(p01), hi32(p01));  pk_p[(c4 >> 1) + 1] = pack_bf16(lo32(p23), hi32(p23));
              pk_ds[c4 >> 1] = pack_bf16(lo32(d01), hi32(d01)); pk_ds[(c4 >> 1) + 1] = pack_bf16(lo32(d23), hi32(d23));
              s[c4 >> 1] = pack_bf16(lo32(g01), hi32(g01));
              s[(c4 >> 1) + 1] = pack_bf16(lo32(g23), hi32(g23));
            }
          } else {
#pragma unroll
            for (int c4 = 0; c4 < 32; c4 += 4) {
              const float4 qc = *reinterpret_cast<const float4 *>(&mq.qcmp[c0 + c4]);
              const ulonglong2 ls = *reinterpret_cast<const ulonglong2 *>(&mq.lse2[c0 + c4]);   // -lse2 pairs
              const ulonglong2 dv = *reinterpret_cast<const ulonglong2 *>(&mq.dvec[c0 + c4]);   // -D pairs
              const uint64_t t01 = ffma2(pk2u(s[c4 + 0], s[c4 + 1]), ksc2, ls.x), t23 = ffma2(pk2u(s[c4 + 2], s[c4 + 3]), ksc2, ls.y);
              const uint64_t p01 = pk2(fast_exp2(ki_j < qc.x ? lo32(t01) : -INFINITY), fast_exp2(ki_j < qc.y ? hi32(t01) : -INFINITY));
              const uint64_t p23 = pk2(fast_exp2(ki_j < qc.z ? lo32(t23) : -INFINITY), fast_exp2(ki_j < qc.w ? hi32(t23) : -INFINITY));
              // keep multipliers m of the four (query, key) pairs: dV takes P∘m, dS = P∘(m∘dP - D)
              uint64_t d01, d23, pm01 = p01, pm23 = p23;
              if (DROPOUT) {
                const uint64_t m01 = pk2((kwb >> (c4 + 0)) & 1u ? kscale : 0.f, (kwb >> (c4 + 1)) & 1u ? kscale : 0.f);
                const uint64_t m23 = pk2((kwb >> (c4 + 2)) & 1u ? kscale : 0.f, (kwb >> (c4 + 3)) & 1u ? kscale : 0.f);
                d01 = fmul2(p01, ffma2(pk2u(dp[c4 + 0], dp[c4 + 1]), m01, dv.x));
                d23 = fmul2(p23, ffma2(pk2u(dp[c4 + 2], dp[c4 + 3]), m23, dv.y));
                pm01 = fmul2(p01, m01); pm23 = fmul2(p23, m23);
              } else {
                d01 = fmul2(p01, fadd2(pk2u(dp[c4 + 0], dp[c4 + 1]), dv.x));
                d23 = fmul2(p23, fadd2(pk2u(dp[c4 + 2], dp[c4 + 3]), dv.y));
              }
              const uint64_t g01 = fmul2(d01, kst2), g23 = fmul2(d23, kst2);
              pk_p[c4 >> 1] = pack_bf16(lo32(pm01), hi32(pm01));  pk_p[(c4 >> 1) + 1] = pack_bf16(lo32(pm23), hi32(pm23));
              pk_ds[c4 >> 1] = pack_bf16(lo32(d01), hi32(d01)); pk_ds[(c4 >> 1) + 1] = pack_bf16(lo32(d23), hi32(d23));
              // reuse s[] as the staging copy (dS * key scale) for dQ
              s[c4 >> 1] = pack_bf16(lo32(g01), hi32(g01));
              s[(c4 >> 1) + 1] = pack_bf16(lo32(g23), hi32(g23));
            }
          }
          tmem_st16(r_st + (cc >> 1), pk_p);
          tmem_st16(r_st + 64 + (cc >> 1), pk_ds);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          sts128(ds_dst + ((static_cast<uint32_t>((cc >> 3) + q) ^ r7) << 4), s[4 * q], s[4 * q + 1], s[4 * q + 2], s[4 * q + 3]);
      }
#ifdef LSH_TRACE_WARP
      if (warp == LSH_TRACE_WARP && lane == 0) BT_TRACE(it.n, 28);
#endif
      tmem_st_wait();
#ifdef LSH_TRACE_WARP
      if (warp == LSH_TRACE_WARP && lane == 0) BT_TRACE(it.n, 29);
#endif
      fence_proxy_async();                                 // dS staging writes -> UMMA (async proxy)
#ifdef LSH_TRACE_WARP
      if (warp == LSH_TRACE_WARP && lane == 0) BT_TRACE(it.n, 30);
#endif
      tc_fence_before();
      mbar_arrive(&sh.pds_full[h]);
      if (lane == 0) BT_TRACE(it.n, 3 + warp);

      if (it.iter_end) {
        // this thread is done with the key tile (and, at a segment end, with the trailing query tile)
        mbar_arrive(&sh.empty[slk]);
        if (it.real && it.last_seg) mbar_arrive(&sh.empty[bt_slot(it.seq_k + 1)]);
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ================================ epilogue warpgroup ================================================
    // Per real iteration (key chunk t): thread j drains dK^ (length-normalisation VJP, App. B5), dQ and dV of
    // token j from TMEM and writes one dq row and one dv row.  Runs concurrently with the next softmax passes.
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    for (BtWalk w(g0, g1, p.n_chunks); w.valid(); w.next()) {
      const BtItem it = w.item();
      if (!it.iter_end) continue;
      const uint32_t slk = bt_slot(it.seq_k);
      if (it.real) {
        mbar_wait(&sh.full[slk], bt_phase(it.seq_k));
        const float ksc_j = sh.meta[slk].kscl[row];
        const int tk = sh.meta[slk].tk[row];
        const int64_t orow = (static_cast<int64_t>(it.u) * p.N + tk) * 64;
        const uint32_t kt_row = tiles_u32 + slk * BT_TILE_BYTES + row * 128, r7 = static_cast<uint32_t>(row & 7);
        mbar_wait(&sh.dq_full[it.rit & 1], (it.rit >> 1) & 1);
        mbar_wait(&sh.kv_full, it.rit & 1);
        tc_fence_after();
        if (row == 0) BT_TRACE(it.n, 25);
        uint32_t dk0[32], dk1[32];
        __nv_bfloat16 *dq_dst = p.dq_out + orow, *dv_dst = p.dv_out + orow;
        tmem_ld32(t_lane + 256, dk0);
        tmem_ld32(t_lane + 288, dk1);
        // dV first: together with the dK^ loads above it empties both accumulators, so the next key chunk's dV / dK^
        // MMAs (kv_free) wait for four TMEM loads instead of the whole epilogue
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t g[32];
          tmem_ld32(t_lane + 320 + 32 * half, g);
          tmem_ld_wait_dep(g);
          if (half == 1) {
            tmem_ld_wait_dep(dk0);
            tmem_ld_wait_dep(dk1);
            tc_fence_before();
            mbar_arrive(&sh.kv_free);
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(g[8 * q4 + 0]), __uint_as_float(g[8 * q4 + 1]));
            v.y = pack_bf16(__uint_as_float(g[8 * q4 + 2]), __uint_as_float(g[8 * q4 + 3]));
            v.z = pack_bf16(__uint_as_float(g[8 * q4 + 4]), __uint_as_float(g[8 * q4 + 5]));
            v.w = pack_bf16(__uint_as_float(g[8 * q4 + 6]), __uint_as_float(g[8 * q4 + 7]));
            *reinterpret_cast<uint4 *>(dv_dst + half * 32 + q4 * 8) = v;
          }
        }
        float dot = 0.f;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint4 raw = lds128u(kt_row + ((static_cast<uint32_t>(ch) ^ r7) << 4));
          const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 q2 = unpack_bf16(rw[e]);
            const int c = ch * 8 + e * 2;
            dot = fmaf(__uint_as_float(c < 32 ? dk0[c] : dk1[c - 32]), q2.x, dot);
            dot = fmaf(__uint_as_float(c + 1 < 32 ? dk0[c + 1] : dk1[c + 1 - 32]), q2.y, dot);
          }
        }
        const float r_j = 0.125f * kLog2e / ksc_j;       // sqrt(mean(q^2) + eps)
        const float a_j = 0.125f / r_j;
        const float c_j = dot * 0.125f / (64.f * r_j * r_j * r_j);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t g[32];
          tmem_ld32(t_lane + 384 + 64 * (it.rit & 1) + 32 * half, g);
          tmem_ld_wait_dep(g);
          if (half == 1) {                               // the dQ slot may be re-used by the next fresh accumulation
            tc_fence_before();
            mbar_arrive(&sh.dq_free[it.rit & 1]);
          }
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int ch = half * 4 + c4;
            const uint4 raw = lds128u(kt_row + ((static_cast<uint32_t>(ch) ^ r7) << 4));
            const uint32_t rw[4] = {raw.x, raw.y, raw.z, raw.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 q2 = unpack_bf16(rw[e]);
              const int c = c4 * 8 + e * 2;              // column within this half
              const float k0 = __uint_as_float(half == 0 ? dk0[c] : dk1[c]), k1 = __uint_as_float(half == 0 ? dk0[c + 1] : dk1[c + 1]);
              o[e] = pack_bf16(__uint_as_float(g[c]) + k0 * a_j - q2.x * c_j, __uint_as_float(g[c + 1]) + k1 * a_j - q2.y * c_j);
            }
            uint4 v; v.x = o[0]; v.y = o[1]; v.z = o[2]; v.w = o[3];
            *reinterpret_cast<uint4 *>(dq_dst + ch * 8) = v;
          }
        }
        // last read of the key tile's rows is behind us: give the ring slot back
        mbar_arrive(&sh.empty[slk]);
        if (it.last_seg) mbar_arrive(&sh.empty[bt_slot(it.seq_k + 1)]);
        if (row == 0) BT_TRACE(it.n, 24);
      } else {
        mbar_arrive(&sh.empty[slk]);                           // pre item: this warpgroup never touches the tile
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 14) tmem_dealloc(tmem, 512);
}

// Round-2 restructuring attempts of this kernel (all measured slower or equal at config 2; the
// code is in the history at commit "Backward kernel: A/B switches ..."): precomputed interval bounds instead of the binary
// search (-0.7 %, but +12 us in chunk_possort_kernel), 16-query blocks with TMEM loads one block ahead (+9 %), early
// release of the key tile's ring slot by the epilogue (+-0).
bool attend_bwd_tc_uses_bounds() { return false; }

int attend_bwd_tc_run(const AttendBwdTcParams &p, int BH, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(BT_NST) * BT_TILE_BYTES + 2 * BT_DS_BYTES + 1024;
  LSH_OPT_IN_SMEM(attend_bwd_tc_kernel<false>);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int total = BH * p.n_chunks;
  int grid = total < sms ? total : sms;
  if (const char *e = getenv("LSH_ATTN_MAX_CTAS")) {   // test hook: force many chunks per CTA on small problems
    const int m = atoi(e);
    if (m > 0 && m < grid) grid = m;
  }
  if (p.keep_bits_t != nullptr) {
    LSH_OPT_IN_SMEM(attend_bwd_tc_kernel<true>);
    attend_bwd_tc_kernel<true><<<grid, BT_THREADS, smem, stream>>>(p, total);
  } else {
    attend_bwd_tc_kernel<false><<<grid, BT_THREADS, smem, stream>>>(p, total);
  }
  LSH_CHECK_LAUNCH("attend_bwd_tc_kernel");
  return 0;
}

}  // namespace lsh
