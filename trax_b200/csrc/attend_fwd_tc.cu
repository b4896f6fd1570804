// Chunked look-back attention, forward, on the 5th-generation tensor cores (tcgen05 + TMEM).
// Same contract as attend_fwd.cu (EA:1958-1986); specialised for chunk_len 128 with a 2-chunk window
// (n_chunks_before + n_chunks_after == 1), dq = dv = 64 — the shape of every long-sequence config.
//
// Persistent, warp-specialised: one CTA per SM walks a contiguous range of (unit, chunk) work items.
//   warps 12-13 producers: position-sorted sticker (requested three tiles ahead) -> cp.async row gathers of the
//                          normalised key qhat (BH, L, 64) and of the value half of the qv row into a ring of chunk tiles
//                          (each tile = 128 rows x (64 qhat + 64 v) bf16, SWIZZLE_128B atoms); completion is signalled by
//                          the copies themselves (cp.async.mbarrier.arrive.noinc).  A tile is loaded ONCE and serves as
//                          "own chunk" for chunk c and as look-back for chunk c+1.  Rank r of the position order goes to
//                          row r (even chunks) or 127 - r (odd chunks).
//   warps 14,15 MMA issuers: scores S[part] = Qhat·Khat[part]^T (tcgen05.mma SS, M128 N128 K16, 4 k-steps; part 0 = look-back
//                          tile keys, part 1 = own tile keys) and O += P[part]·V[part] (tcgen05.mma TS, P from TMEM, V
//                          MN-major, 8 k-steps), parts issued skewed by one chunk; completion and tile release via
//                          tcgen05.commit -> mbarriers.
//   warps 0-3, 4-7 two softmax warpgroups, warpgroup p owns part p of EVERY chunk: thread = query row = TMEM lane.
//                          One pass: t = s * a_i - m_i (per-row scale and shift: q_i·k_j = 8 r_i (qhat_i·qhat_j)), p = exp2(t),
//                          P (bf16) into its own TMEM buffer; the causal / self masks are an INTERVAL of columns per row
//                          (position-sorted tiles): 32-column blocks no row of the warp sees are skipped, the others go
//                          through one block routine with a per-row visibility bit mask.  The score buffer is handed back to the
//                          issuer by the pass itself.
//   warps 8-11 epilogue warpgroup: O (TMEM) * 1/l -> bf16 row -> ticker slot; l = l_part0 + l_part1 through shared memory.
// The softmax shift m_i is the un-masked self score (it bounds every score of the row, Cauchy-Schwarz with |khat| ~ 1),
// so no max pass is needed; a row whose only visible keys are itself / its copy (EA "-1e5" class) keeps exactly those.
// TMEM map and the per-block code are documented at TcShared / softmax_block_*; DESIGN.md section 4.3 has the measurements.
#include "attend_params.cuh"
#include "tc_common.cuh"

#include <stdlib.h>

// A/B switches (round 2): TC_FWD_BOUNDS 1 = look-back interval from the bounds chunk_possort_kernel precomputes, 0 = 8-step
// binary search per row and pass (round 1); TC_FWD_REDO 1 = rows whose sum under/overflows are queued for the exact redo.
// Measured at config 2 (ncu, one box, interleaved): search + redo 288.5 us, bounds + redo 298.5 us, search without redo 287 us.
#ifndef TC_FWD_BOUNDS
#define TC_FWD_BOUNDS 0
#endif
#ifndef TC_FWD_REDO
#define TC_FWD_REDO 1
#endif
// TC_FWD_TMA 1: the producers fetch the tile rows with TMA (cp.async.bulk.tensor ... tile::gather4: four gathered rows per
// instruction, written by the async proxy straight into the SWIZZLE_128B atoms, completion counted in bytes on the tile's
// mbarrier) — 2 instructions per lane and tile; 0: 32 cp.async (16 bytes each) per lane and tile (round 1).
// Measured at config 2 (ncu, interleaved on one box, both parity-green over the whole GPU suite): cp.async 290.8-293.9 us,
// TMA gather4 298.8-305.2 us (second box: 268-272 vs 278-280 us) — the row gather is 3.5 % SLOWER through the TMA unit
// although the producers issue 16x fewer instructions, so cp.async stays the default and the TMA path the A/B variant.
#ifndef TC_FWD_TMA
#define TC_FWD_TMA 0
#endif

namespace lsh {

constexpr int TC_C = 128;
constexpr int TC_NST = 6;                    // tile ring depth
constexpr int TC_THREADS = 512;              // 8 softmax + 4 epilogue + 2 producer + 2 issuer warps
constexpr int TC_TILE_BYTES = 2 * TC_C * 128;   // K rows then V rows
constexpr uint32_t TC_IDESC_S = make_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
constexpr uint32_t TC_IDESC_O = make_idesc_bf16(128, 64, 0, 1);    // P (TMEM)    x V (MN-major)

struct __align__(16) TcTileMeta {
  float kinfo[TC_C];    // kv_info (+1 applied; negative = padding), fp32 like EA:148-149 (generic path)
  int pos[TC_C];        // 0-based positions; ascending along the tile's rank order (see chunk_possort_kernel)
  int tk[TC_C];         // ticker values
  int bnd[TC_C];        // neighbour-chunk interval bounds of the row (chunk_possort_kernel)
  float2 am[TC_C];      // {a = 8 r log2e, m2 = a |qhat|^2} of the row's token (query-side scale and softmax shift)
  float vmin[2], vmax[2];   // min / max of the valid kinfo per producer warp (visibility test, generic path)
};

struct __align__(16) TcShared {
  TcTileMeta meta[TC_NST];
  uint64_t full[TC_NST], empty[TC_NST];
  uint64_t s_full[2], s_free[2], p_full[2], p_free[2];   // per window part: score buffer / probability buffer hand-offs
  uint64_t o_full[2], o_free[2], l_full[2];              // per output accumulator (64 TMEM columns each) / row statistics
  float row_l[2][2][TC_C], row_m2[2][TC_C], row_off[2][TC_C];   // softmax -> epilogue hand-off per output buffer
  int row_tk[2][TC_C];
  uint32_t tmem_base;
};

// TMEM map (512 columns).  A chunk's 256 keys are scored as two independent halves (part 0: first window tile, part 1:
// second), each owned by one softmax warpgroup with its own buffers: scores S[part] at 128 * part (fp32, 128 columns),
// probabilities P[part] at 256 + 64 * part (bf16 pairs, 64 columns), two output accumulators O at 384 and 448.
// P is NOT written over S: a score buffer is free for the next chunk's score MMA the moment its softmax pass has read
// it — neither the PV MMA nor the epilogue sits between two passes of a warpgroup.
constexpr uint32_t TC_P_COL = 256;
constexpr uint32_t TC_O_COL = 384;

// Enumerates the work items of one CTA and the tile sequence numbers they use.
struct Walker {
  int g, g_end, n_chunks, k, n;     // global chunk id, end, chunks per unit, index in range, seq of the 2nd window tile
  int u, c;
  bool reuse;
  __device__ Walker(int g0, int g1, int nc) : g(g0), g_end(g1), n_chunks(nc), k(0), n(1) {
    u = g / nc; c = g - u * nc; reuse = false;
  }
  __device__ bool valid() const { return g < g_end; }
  __device__ bool next_reuses() const { return (g + 1 < g_end) && (c + 1 < n_chunks); }
  __device__ void next() {
    ++g; ++k; ++c;
    if (c == n_chunks) { c = 0; ++u; reuse = false; } else { reuse = true; }
    n += reuse ? 1 : 2;
  }
};

__device__ __forceinline__ uint32_t slot_of(int n) { return static_cast<uint32_t>(n % TC_NST); }
__device__ __forceinline__ uint32_t phase_of(int n) { return static_cast<uint32_t>((n / TC_NST) & 1); }

// trace slots per chunk: 0 S issued, 1 PV issued, 2 s_full seen, 3 pass done, 4 o_full seen, 5 epilogue done, 6 tile issued, 7 tile landed
// (compiled in only with -DLSH_TRACE: the stamps cost instruction-cache space in every role)
#ifdef LSH_TRACE
#define TC_TRACE(k, slot) do { if (p.trace && blockIdx.x == 0 && (k) < 120) p.trace[(k) * 16 + (slot)] = clock64(); } while (0)
#define TC_TRACE_ON 1
#else
#define TC_TRACE(k, slot) do { } while (0)
#define TC_TRACE_ON 0
#endif

constexpr float kBig = 1e9f * kLog2e, kSelf = 1e5f * kLog2e;   // EA:152-159 masks, log2 domain
constexpr float kRedoBelow = 1e-30f, kRedoAbove = 1e30f;       // row sums outside this range are redone with the true max

// 32 score columns of one query row: t = s * a_i - m_i, p = 2^t, packed to bf16 into the P buffer.  Scale and shift are
// per-ROW registers (no loads); bit c of `vis` says whether column c is visible.  ONE code copy serves interior blocks
// (all ones) and boundary blocks: fewer instructions for the former were measured to be worth less than the
// instruction-cache footprint of separate variants (A/B on one box: 0.329 vs 0.337 ms per launch + auxiliaries).
// 2^x on the FMA / ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max relative error 7.5e-5 — P is rounded to
// bf16, 4e-3, right after): every TC_FWD_POLY-th column pair of a block takes this instead of MUFU.EX2, which two softmax
// warps of one sub-partition otherwise keep ~85 % busy.  0 = off.
#ifndef TC_FWD_POLY
#define TC_FWD_POLY 0
#endif
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.f);                                    // (masked / far-away scores: 2^-125, rounds to 0 in bf16)
  const float y = x + 12582912.f;                          // 1.5 * 2^23: round-to-nearest integer part in the low mantissa bits
  const float f = x - (y - 12582912.f);                    // in [-0.5, 0.5]
  float p = fmaf(f, 0.05517132f, 0.24261054f);
  p = fmaf(p, f, 0.69326097f);
  p = fmaf(p, f, 0.9999281f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(y) << 23));
}
__device__ __forceinline__ void softmax_block_mask(const uint32_t (&r)[32], uint64_t a2, uint64_t mm2, uint32_t vis, uint32_t t_dst,
                                                   uint64_t &l2) {
  uint32_t pk[16];
#pragma unroll
  for (int c2 = 0; c2 < 32; c2 += 2) {
    const uint64_t t = ffma2(pk2u(r[c2], r[c2 + 1]), a2, mm2);
    if (TC_FWD_POLY > 0 && ((c2 >> 1) % (TC_FWD_POLY > 0 ? TC_FWD_POLY : 1)) == 0) {
      const float q0 = poly_exp2(lo32(t)), q1 = poly_exp2(hi32(t));
      const float p0 = (vis & (1u << c2)) ? q0 : 0.f, p1 = (vis & (2u << c2)) ? q1 : 0.f;
      l2 = fadd2(l2, pk2(p0, p1));
      pk[c2 >> 1] = pack_bf16(p0, p1);
      continue;
    }
    const float p0 = fast_exp2((vis & (1u << c2)) ? lo32(t) : -INFINITY);
    const float p1 = fast_exp2((vis & (2u << c2)) ? hi32(t) : -INFINITY);
    l2 = fadd2(l2, pk2(p0, p1));
    pk[c2 >> 1] = pack_bf16(p0, p1);
  }
  tmem_st16(t_dst, pk);
}
// Generic path (non-causal, padding mask, other windows): the reference's subtractive masks in order (EA:150-159) with the
// key kv_info read from shared memory.
// `keep`: attention-dropout bits of the block's 32 columns (EA:254-262; all ones without dropout) — applied to P after the
// row sum, which, like the reference's log-sum-exp, does not see the dropout.
__device__ __forceinline__ void softmax_block_generic(const uint32_t (&r)[32], const float *kin, float qi, float a, float m2, int causal,
                                                      int masked, uint32_t keep, uint32_t t_dst, float &l) {
  uint32_t pk[16];
#pragma unroll
  for (int c4 = 0; c4 < 32; c4 += 4) {
    const float4 ki4 = *reinterpret_cast<const float4 *>(kin + c4);
    const float kis[4] = {ki4.x, ki4.y, ki4.z, ki4.w};
    float pv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float t = fmaf(__uint_as_float(r[c4 + e]), a, -m2);
      if (causal && qi < kis[e]) t -= kBig;
      if (qi == kis[e]) t -= kSelf;
      if (masked && kis[e] < 0.f) t -= kBig;
      pv[e] = fast_exp2(t);
    }
    l += (pv[0] + pv[1]) + (pv[2] + pv[3]);
    const uint32_t kq = keep >> c4;
    pk[c4 >> 1] = pack_bf16((kq & 1u) ? pv[0] : 0.f, (kq & 2u) ? pv[1] : 0.f);
    pk[(c4 >> 1) + 1] = pack_bf16((kq & 4u) ? pv[2] : 0.f, (kq & 8u) ? pv[3] : 0.f);
  }
  tmem_st16(t_dst, pk);
}

// Slot-ordered tiles, causal, no padding mask (the attention-dropout calls of the long-sequence configs): ONE compare per
// element — key position + 1 against qcmp = q_info, or q_info + 0.5 for a row that sees no earlier key and therefore keeps
// exactly its own "-1e5" class (EA:150-155; the caller moves that -1e5 into the reported log-sum-exp) — packed math as in
// the position-sorted routine, and the keep bits applied to P after the row sum.
__device__ __forceinline__ void softmax_block_causal_keep(const uint32_t (&r)[32], const float *kin, float qcmp, uint64_t a2,
                                                          uint64_t mm2, uint32_t keep, uint32_t t_dst, uint64_t &l2) {
  uint32_t pk[16];
#pragma unroll
  for (int c4 = 0; c4 < 32; c4 += 4) {
    const float4 ki = *reinterpret_cast<const float4 *>(kin + c4);
    const uint64_t t01 = ffma2(pk2u(r[c4], r[c4 + 1]), a2, mm2), t23 = ffma2(pk2u(r[c4 + 2], r[c4 + 3]), a2, mm2);
    const float p0 = fast_exp2(ki.x < qcmp ? lo32(t01) : -INFINITY), p1 = fast_exp2(ki.y < qcmp ? hi32(t01) : -INFINITY);
    const float p2 = fast_exp2(ki.z < qcmp ? lo32(t23) : -INFINITY), p3 = fast_exp2(ki.w < qcmp ? hi32(t23) : -INFINITY);
    l2 = fadd2(l2, fadd2(pk2(p0, p1), pk2(p2, p3)));
    const uint32_t kq = keep >> c4;
    pk[c4 >> 1] = pack_bf16((kq & 1u) ? p0 : 0.f, (kq & 2u) ? p1 : 0.f);
    pk[(c4 >> 1) + 1] = pack_bf16((kq & 4u) ? p2 : 0.f, (kq & 8u) ? p3 : 0.f);
  }
  tmem_st16(t_dst, pk);
}

// SORTED: causal, no padding mask, look-back window (every long-sequence config) — interval masks on position-sorted tiles.
// !SORTED: the reference's masks evaluated per element from the keys' kv_info.  Two instantiations keep each one's code
// (instruction-cache footprint) small.
template <bool SORTED>
__global__ void __launch_bounds__(TC_THREADS, 1) attend_fwd_tc_kernel(const __grid_constant__ AttendFwdParams p, int total_chunks) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *tiles = smem;                                    // [TC_NST][K 16 KB | V 16 KB]
  __shared__ TcShared sh;                                   // static: keeps metadata accesses in the shared space (LDS)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long t_cta_start = TC_TRACE_ON ? clock64() : 0;
  // contiguous, balanced range of chunks for this CTA
  const int g0 = static_cast<int>(static_cast<int64_t>(total_chunks) * blockIdx.x / gridDim.x);
  const int g1 = static_cast<int>(static_cast<int64_t>(total_chunks) * (blockIdx.x + 1) / gridDim.x);

  if (warp == 14) tmem_alloc(&sh.tmem_base, 512);
#if TC_FWD_TMA
  if (tid == 12 * 32) { tma_prefetch_desc(&p.tm_k); tma_prefetch_desc(&p.tm_v); }
#endif
  if (tid == 0) {
    // full: per producer thread one arrival by its copies (cp.async ... noinc) and one ordinary arrival that releases its
    // plain metadata stores
    for (int i = 0; i < TC_NST; ++i) { mbar_init(&sh.full[i], 128); mbar_init(&sh.empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh.s_full[i], 1); mbar_init(&sh.p_free[i], 1);     // tcgen05.commit arrivals
      mbar_init(&sh.s_free[i], 4); mbar_init(&sh.p_full[i], 4);     // one arrival per softmax warp of the part's warpgroup
      mbar_init(&sh.o_full[i], 1); mbar_init(&sh.o_free[i], 4);     // one arrival per epilogue warp
      mbar_init(&sh.l_full[i], 256);                                 // every softmax thread releases its own row statistics
    }
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;
  const uint32_t tiles_u32 = smem_u32(tiles);

  if (warp == 12 || warp == 13) {
    // ================================ producers ======================================================
    const int pw = warp & 1;                              // rows [64*pw, 64*pw + 64) of the tile
    // tile stream of this CTA in sequence order: (seq, unit, chunk)
    Walker wk(g0, g1, p.n_chunks);
    bool first_done = false;                              // the extra look-back tile of a non-reusing chunk
    auto next_tile = [&](int &n, int &u, int &cc) -> bool {
      while (wk.valid()) {
        const int c0 = ((wk.c - p.nb) % p.n_chunks + p.n_chunks) % p.n_chunks;   // first window chunk (EA:137-141)
        if (!wk.reuse && !first_done) { first_done = true; n = wk.n - 1; u = wk.u; cc = c0; return true; }
        n = wk.n; u = wk.u; cc = (c0 + 1) % p.n_chunks;
        wk.next(); first_done = false;
        return true;
      }
      return false;
    };
    // rank r of the position-sorted chunk goes to tile row r (even chunks) or 127 - r (odd chunks): the two softmax
    // warpgroups own alternate chunks, so their heavy warps (the rows that see the most keys) sit on
    // different SM sub-partitions.
    auto fetch_sticker = [&](int u, int cc, int &tka, int &tkb, int &bda, int &bdb) {
      const int64_t off = static_cast<int64_t>(u) * p.N + cc * TC_C + 64 * pw;
      // attention dropout indexes its keep matrix by SLOT: those calls keep the reference's order inside a chunk
      const int32_t *stk = ((!SORTED && p.keep_bits) ? p.sticker : p.sticker2) + off;
      tka = __ldg(stk + lane); tkb = __ldg(stk + 32 + lane);
#if TC_FWD_BOUNDS
      if constexpr (SORTED) { bda = __ldg(p.bounds + off + lane); bdb = __ldg(p.bounds + off + 32 + lane); }
#endif
    };
    // Requests run three tiles ahead of the copies so that the dependent chain sticker -> position -> row address never
    // exposes a global-load latency (three statically named request slots: no register rotation, no early scoreboard wait).
    struct TileReq { int n, u, cc, tka, tkb, bda, bdb; bool have; };
    auto request = [&](TileReq &r) {
      r.tka = 0; r.tkb = 0; r.bda = 0; r.bdb = 0;
      r.have = next_tile(r.n, r.u, r.cc);
      if (r.have) fetch_sticker(r.u, r.cc, r.tka, r.tkb, r.bda, r.bdb);
    };
    constexpr bool sorted_path = SORTED;
    [[maybe_unused]] const int ch = lane & 15, hi = lane >> 4;   // 16-byte piece of the 256-byte row pair; row parity
    auto issue = [&](const TileReq &r) {
      const int n = r.n, u = r.u, tka = r.tka, tkb = r.tkb;
      const uint32_t slot = slot_of(n);
      const int flip = (r.cc & 1) ? 127 : 0;                // row = rank ^ flip  (127 - rank for odd chunks)
      mbar_wait<64>(&sh.empty[slot], phase_of(n) ^ 1);
      const int b = u / p.H, h = u - b * p.H;
      const int pa = tka % p.L, pb = tkb % p.L;
      TcTileMeta &mt = sh.meta[slot];
      const int rowa = (64 * pw + lane) ^ flip, rowb = (64 * pw + 32 + lane) ^ flip;
      mt.pos[rowa] = pa;    mt.pos[rowb] = pb;
      mt.tk[rowa] = tka;    mt.tk[rowb] = tkb;
#if TC_FWD_BOUNDS
      if constexpr (sorted_path) { mt.bnd[rowa] = r.bda; mt.bnd[rowb] = r.bdb; }
#endif
      const float2 *rm = p.rowmeta + static_cast<int64_t>(u) * p.L;
      cp_async8(smem_u32(&mt.am[rowa]), rm + pa);
      cp_async8(smem_u32(&mt.am[rowb]), rm + pb);
      if constexpr (!sorted_path) {                         // generic path: kv_info and the window visibility range
        bool va = true, vb = true;
        if (p.masked) {
          va = p.mask[static_cast<int64_t>(b) * p.L + pa] != 0;
          vb = p.mask[static_cast<int64_t>(b) * p.L + pb] != 0;
        }
        const float kia = static_cast<float>((va ? pa : -pa) + 1), kib = static_cast<float>((vb ? pb : -pb) + 1);
        mt.kinfo[rowa] = kia; mt.kinfo[rowb] = kib;
        float mn = fminf(kia > 0.f ? kia : INFINITY, kib > 0.f ? kib : INFINITY);
        float mx = fmaxf(kia > 0.f ? kia : -INFINITY, kib > 0.f ? kib : -INFINITY);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) { mt.vmin[pw] = mn; mt.vmax[pw] = mx; }
      }
#if TC_FWD_TMA
      // TMA row gather: lanes 0-15 fetch the normalised keys qhat (BH, L, 64), lanes 16-31 the value halves of the qv rows,
      // each lane the four ranks 4 q .. 4 q + 3 of this warp's 64 (q = 16 pw + lane % 16).  Rank r lives in tile row r ^ flip,
      // so the four land in the aligned row group 4 G (G = q, or 31 - q for a flipped tile) — in reverse order when flipped.
      {
        const int l16 = lane & 15, isv = lane >> 4;
        int pr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                           // rank 4 l16 + j of the warp's 64: pa of that lane, or pb of lane - 32
          const int va = __shfl_sync(0xffffffffu, pa, (4 * l16 + j) & 31), vb = __shfl_sync(0xffffffffu, pb, (4 * l16 + j) & 31);
          pr[j] = l16 < 8 ? va : vb;
        }
        const int q = 16 * pw + l16, G = flip ? 31 - q : q;
        const int i0 = flip ? pr[3] : pr[0], i1 = flip ? pr[2] : pr[1], i2 = flip ? pr[1] : pr[2], i3 = flip ? pr[0] : pr[3];
        const uint32_t dst = tiles_u32 + slot * TC_TILE_BYTES + (isv ? TC_C * 128 : 0) + G * 512;
        if (isv) {
          const int rb = (b * p.L) * p.H + h;                   // row of (b, position 0, h) in the (B*L*H, 128) view of qv
          tma_gather4(dst, &p.tm_v, &sh.full[slot], 64, rb + i0 * p.H, rb + i1 * p.H, rb + i2 * p.H, rb + i3 * p.H);
        } else {
          const int rb = u * p.L;
          tma_gather4(dst, &p.tm_k, &sh.full[slot], 0, rb + i0, rb + i1, rb + i2, rb + i3);
        }
      }
      // completion: the copies' bytes (16 KB per warp, expected by lane 0's arrival) + the rowmeta cp.async + one arrival
      // per thread releasing its plain metadata stores
      cp_async_mbar_arrive_noinc(&sh.full[slot]);
      if (lane == 0) mbar_arrive_expect_tx(&sh.full[slot], 64 * 256);
      else mbar_arrive(&sh.full[slot]);
#else
      // 16-byte pieces 0-7 of a row: normalised key qhat (BH, L, 64); pieces 8-15: value, second half of the qv row.
      // Destination of rank R = 64 pw + 2 i + hi, piece c: tile + (X << 7) + (((c ^ X) & 7) << 4) with X = R ^ flip.  All
      // bit fields are disjoint, so this is  tile + (lane_tile_const ^ imm(i)):  one LOP3 + one IADD per copy.
      const char *base = ch < 8 ? reinterpret_cast<const char *>(p.qhat + static_cast<int64_t>(u) * p.L * 64 + ch * 8)
                                : reinterpret_cast<const char *>(p.qv + (static_cast<int64_t>(b) * p.L * p.H + h) * 128 + ch * 8);
      const uint32_t rbytes = ch < 8 ? 128u : static_cast<uint32_t>(p.H) * 256u;
      const uint32_t tile = tiles_u32 + slot * TC_TILE_BYTES + (ch < 8 ? 0 : TC_C * 128);
      const uint32_t lane_const = (static_cast<uint32_t>((64 * pw + hi) ^ flip) << 7) | (static_cast<uint32_t>((ch ^ flip ^ hi) & 7) << 4);
#pragma unroll 1
      for (int io = 0; io < 8; ++io) {                     // 4 copies per trip: compact code, constants stay immediates
        const int psel = io < 4 ? pa : pb;
        const uint32_t dst = tile + (lane_const ^ (static_cast<uint32_t>(io) << 10));
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int pr = __shfl_sync(0xffffffffu, psel, (8 * io + 2 * ii + hi) & 31);
          cp_async16(dst ^ ((static_cast<uint32_t>(ii) << 8) ^ (static_cast<uint32_t>(ii) << 5)), base + static_cast<uint64_t>(static_cast<uint32_t>(pr)) * rbytes);
        }
      }
      // Completion is signalled by the copies themselves (no wait here): every free ring slot is a tile in flight.
      // The ordinary arrival (release) publishes this thread's plain metadata stores.
      cp_async_mbar_arrive_noinc(&sh.full[slot]);
      mbar_arrive(&sh.full[slot]);
#endif
    };
    TileReq r0, r1, r2;
    request(r0); request(r1); request(r2);
    for (;;) {
      if (!r0.have) break;
      issue(r0); request(r0);
      if (!r1.have) break;
      issue(r1); request(r1);
      if (!r2.have) break;
      issue(r2); request(r2);
    }
  } else if (warp == 14) {
    // ================================ S issuer ========================================================
    // Part p of chunk k (part 0: the first window tile's keys, part 1: the second's) goes out as soon as its tiles have
    // landed and warpgroup p has read the previous chunk's scores out of the buffer.  The warp runs converged
    // (warp-uniform values); one elected lane issues.  Descriptors: constant hi, lo = base + (offset >> 4).
    // Issue order is skewed by one chunk between the parts — step i: (chunk i, part 0), then (chunk i-1, part 1) — so the
    // two warpgroups work on different chunks at any time: consecutive chunks have opposite row orders, hence their heavy
    // warps (the rows that see the most keys) sit on different SM sub-partitions.
    constexpr uint32_t HI = desc_hi(1024);
    auto issue_s = [&](int k, int n, int part) {
      const uint32_t k0 = tiles_u32 + slot_of(n - 1) * TC_TILE_BYTES, k1 = tiles_u32 + slot_of(n) * TC_TILE_BYTES;
      const uint32_t qa = desc_lo(p.nb ? k1 : k0, 16);
      mbar_wait(&sh.s_free[part], (k & 1) ^ 1);
      if (lane == 0) TC_TRACE(k, part == 0 ? 7 : 12);
      // no proxy fence: the tile's mbarrier phase completes when its cp.async copies have landed, which is what the UMMA
      // operand reads are ordered after (same protocol as CUTLASS's sm100 cp.async mainloop)
      tc_fence_after();
      const uint32_t kb = desc_lo(part ? k1 : k0, 16), s_t = tmem + part * 128;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_ss2(s_t, qa + ks * 2, HI, kb + ks * 2, HI, TC_IDESC_S, ks > 0);
        umma_commit(&sh.s_full[part]);
      }
      __syncwarp();
      if (lane == 0 && part == 1) TC_TRACE(k, 0);
    };
    Walker ws(g0, g1, p.n_chunks);
    bool have_prev = false;
    int pk = 0, pn = 0;
    while (ws.valid() || have_prev) {
      const bool cur = ws.valid();
      if (cur) {
        mbar_wait(&sh.full[slot_of(ws.n - 1)], phase_of(ws.n - 1));
        mbar_wait(&sh.full[slot_of(ws.n)], phase_of(ws.n));
        issue_s(ws.k, ws.n, 0);
      }
      if (have_prev) issue_s(pk, pn, 1);
      have_prev = cur;
      if (cur) { pk = ws.k; pn = ws.n; ws.next(); }
    }
  } else if (warp == 15) {
    // ================================ PV issuer =======================================================
    // A second issuing warp so that a PV never queues behind a score MMA that is still waiting for tiles.  Same skewed
    // order as the score issuer.
    constexpr uint32_t HI = desc_hi(1024);
    auto issue_pv = [&](int k, int n, int part, bool release_first) {
      const uint32_t ob = k & 1, jo = k >> 1;
      const uint32_t o_t = tmem + TC_O_COL + ob * 64;
      const uint32_t vt = desc_lo(tiles_u32 + slot_of(part ? n : n - 1) * TC_TILE_BYTES + TC_C * 128, 1024);
      const uint32_t p_t = tmem + TC_P_COL + part * 64;
      mbar_wait(&sh.p_full[part], k & 1);
      if (lane == 0) TC_TRACE(k, 8 + 2 * part);
      if (part == 0) mbar_wait(&sh.o_free[ob], (jo & 1) ^ 1);       // the epilogue two chunks back has drained this O
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int i = 0; i < 8; ++i) umma_ts2(o_t, p_t + i * 8, vt + i * 128, HI, TC_IDESC_O, (part > 0 || i > 0) ? 1u : 0u);
        umma_commit(&sh.p_free[part]);
        // Tile release.  The score MMAs that read a tile (other issuer) finished before the P of their part existed.
        // Part 1 of chunk k is the last reader of tile n: part 0 of chunk k+1 (its look-back use) was issued just before.
        // A first window tile that is not shared with the previous chunk (range / unit start) has part 0 as only reader.
        if (part == 0) {
          if (release_first) umma_commit(&sh.empty[slot_of(n - 1)]);
        } else {
          umma_commit(&sh.o_full[ob]);
          umma_commit(&sh.empty[slot_of(n)]);
        }
      }
      __syncwarp();
      if (lane == 0) TC_TRACE(k, 9 + 2 * part);
      if (lane == 0 && part == 1) TC_TRACE(k, 1);
    };
    Walker wo(g0, g1, p.n_chunks);
    bool have_prev = false;
    int pk = 0, pn = 0;
    while (wo.valid() || have_prev) {
      const bool cur = wo.valid();
      if (cur) issue_pv(wo.k, wo.n, 0, !wo.reuse);
      if (have_prev) issue_pv(pk, pn, 1, false);
      have_prev = cur;
      if (cur) { pk = wo.k; pn = wo.n; wo.next(); }
    }
  } else if (warp < 8) {
    // ================================ softmax warpgroups ==============================================
    // Column split: warpgroup 0 owns the look-back tile's 128 score columns of EVERY chunk, warpgroup 1 the own tile's.
    // Chunks alternate between the two TMEM regions; a warp that finishes its (position-dependent) share early moves on
    // to the next chunk, where the row order is reversed and it is the heavy one.
    const uint32_t wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;                // query row == TMEM lane (lane quarter = warp id % 4)
    const uint32_t t_wg = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    constexpr bool sorted = SORTED;
    for (Walker wk(g0, g1, p.n_chunks); wk.valid(); wk.next()) {
      const int n = wk.n;
      const uint32_t ob = wk.k & 1, jo = wk.k >> 1;
      const uint32_t t_lane = t_wg + wg * 128;                       // my part's scores
      const uint32_t t_p = t_wg + TC_P_COL + wg * 64;                // my part's probabilities
      const uint32_t sl0 = slot_of(n - 1), sl1 = slot_of(n);
      mbar_wait(&sh.full[sl0], phase_of(n - 1));
      mbar_wait(&sh.full[sl1], phase_of(n));
      const TcTileMeta &m0 = sh.meta[sl0], &m1 = sh.meta[sl1];
      const TcTileMeta &mq = p.nb ? m1 : m0;
      const TcTileMeta &mk = wg ? m1 : m0;                          // the tile whose keys this warpgroup scores
      const int mypos = mq.pos[row];
      const int my_tk = mq.tk[row];                                 // (read here: a load in the hand-off tail sits behind the P stores)
      const float qi = static_cast<float>(mypos + 1);               // q_info = pos + 1 (EA:201)
      const float2 am = mq.am[row];                                  // query-side scale a_i, self score m_i (log2 domain)
      const float a_i = am.x;
      float m2 = am.y, lse_off = 0.f;
      [[maybe_unused]] float qcmp = qi;                              // (slot-ordered causal routine)
      uint32_t need = 0xfu;                                          // 32-column blocks of my tile any row of this warp sees
      int lo = 0, hi = 128;                                          // visible column interval in my tile
      uint32_t keep4[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};   // dropout keep bits per 32-column block
      if constexpr (sorted) {
        // Both tiles are ordered by position (rank r at row r ^ flip), so "key position < query position" (EA:150-152 and the
        // self mask EA:153-155, whose -1e5 entries underflow to exactly 0 next to any visible key) is an interval of columns.
        const int c_lb = wk.c > 0 ? wk.c - 1 : p.n_chunks - 1;      // cyclic look-back (EA:137-141)
        const int flip_own = (wk.c & 1) ? 127 : 0, flip_lb = (c_lb & 1) ? 127 : 0;
        const int myrank = row ^ flip_own;
        // Where my position falls inside the look-back chunk comes precomputed (chunk_possort_kernel): cnt = look-back keys
        // with position < mine, eq = the look-back chunk holds my own position (my copy from the previous hash round).
        // A row without any visible key keeps exactly its "-1e5" class (itself, and that copy), and the -1e5 goes back
        // into the reported log-sum-exp.
#if TC_FWD_BOUNDS
        const int bnd = mq.bnd[row], cnt_lb = bnd & 0xff, eq_lb = (bnd >> 8) & 1;
        const bool lonely = myrank == 0 && cnt_lb == 0;
#else
        const int lb_min = m0.pos[flip_lb];                          // smallest look-back position
        const bool lonely = myrank == 0 && lb_min >= mypos;
#endif
        if (lonely) lse_off = -1e5f;
        if (wg == 0) {
#if TC_FWD_BOUNDS
          const int bound = lonely ? eq_lb : cnt_lb;
#else
          int blo = 0, bhi = 128;                                    // look-back keys with position < mypos (lower bound)
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int mid = (blo + bhi) >> 1;
            const int v = m0.pos[(mid & 127) ^ flip_lb];
            const bool go = blo < bhi;
            if (go && v < mypos) blo = mid + 1;
            else if (go) bhi = mid;
          }
          const int bound = lonely ? (lb_min == mypos ? 1 : 0) : blo;
#endif
          if (flip_lb) { lo = 128 - bound; hi = 128; } else { lo = 0; hi = bound; }
        } else {
          const int self_incl = lonely ? 1 : 0;
          if (flip_own) { lo = row + 1 - self_incl; hi = 128; } else { lo = 0; hi = row + self_incl; }
        }
        const int lo_min = __reduce_min_sync(0xffffffffu, lo), hi_max = __reduce_max_sync(0xffffffffu, hi);
        need = 0u;
#pragma unroll
        for (int bq = 0; bq < 4; ++bq) {
          const int c0 = 32 * bq, c1 = c0 + 32;
          if (!(hi_max <= c0 || lo_min >= c1)) need |= 1u << bq;
        }
      } else {
        if (p.keep_bits) {
          // my row of the (C, W) keep matrix, the 128 columns of my window part: tile rows are slot ^ flip of their chunk
          const int c_k = ((wk.c + static_cast<int>(wg) - p.nb) % p.n_chunks + p.n_chunks) % p.n_chunks;
          const int qslot = row ^ ((wk.c & 1) ? 127 : 0);
          const uint4 kw = __ldg(reinterpret_cast<const uint4 *>(p.keep_bits + qslot * 8 + wg * 4));
          if (c_k & 1) { keep4[0] = __brev(kw.w); keep4[1] = __brev(kw.z); keep4[2] = __brev(kw.y); keep4[3] = __brev(kw.x); }
          else { keep4[0] = kw.x; keep4[1] = kw.y; keep4[2] = kw.z; keep4[3] = kw.w; }
        }
        const float own_ki = mq.kinfo[row];
        const float wmin = fminf(fminf(m0.vmin[0], m0.vmin[1]), fminf(m1.vmin[0], m1.vmin[1]));
        const float wmax = fmaxf(fmaxf(m0.vmax[0], m0.vmax[1]), fmaxf(m1.vmax[0], m1.vmax[1]));
        const bool visible = p.causal ? (wmin < qi) : !(wmin == qi && wmax == qi);
        if (p.causal && !p.masked) {
          // fast causal routine (softmax_block_causal_keep): a row without any earlier key keeps exactly its "-1e5" class
          qcmp = visible ? qi : qi + 0.5f;
          if (!visible) lse_off = -1e5f;
        } else {
          if (!visible) m2 -= kSelf;                                  // only the "-1e5" class is left: shift by it
          if (p.masked && own_ki < 0.f) m2 = -kBig;                   // padding query: any finite result
        }
      }
      const uint64_t a2 = pk2(a_i, a_i), mm2 = pk2(-m2, -m2);
      if (warp == 0 && lane == 0) TC_TRACE(wk.k, 6);
      mbar_wait(&sh.s_full[wg], wk.k & 1);
      mbar_wait(&sh.p_free[wg], (wk.k & 1) ^ 1);                     // the previous chunk's PV has read my P buffer
      tc_fence_after();
      if (warp == 0 && lane == 0) TC_TRACE(wk.k, 2);
      if (warp == 4 && lane == 0) TC_TRACE(wk.k, 13);
      long long *tr2 = (TC_TRACE_ON && p.trace && blockIdx.x == 0 && wk.k < 120 && wg == 0) ? p.trace + 120 * 16 + 148 + wk.k * 12 + (warp & 3) * 3 : nullptr;
      if (tr2 && lane == 0) { tr2[0] = clock64(); tr2[2] = __popc(need) * 16; }
      float l = 0.f;
      uint64_t l2 = 0ull;
      uint32_t ra[32], rb[32];
      // Loads run one needed block ahead; skipped blocks get zeros.
      auto process = [&](const uint32_t (&r)[32], int bq) {
        if constexpr (sorted) {
          const int lr = lo - bq * 32, hr = hi - bq * 32;     // my visible interval [lo, hi) relative to the block
          const uint32_t below_hi = hr >= 32 ? 0xffffffffu : (hr <= 0 ? 0u : ((1u << hr) - 1u));
          const uint32_t below_lo = lr >= 32 ? 0xffffffffu : (lr <= 0 ? 0u : ((1u << lr) - 1u));
          softmax_block_mask(r, a2, mm2, below_hi & ~below_lo, t_p + bq * 16, l2);
        } else {
          const uint32_t kpw = bq == 0 ? keep4[0] : (bq == 1 ? keep4[1] : (bq == 2 ? keep4[2] : keep4[3]));
          if (p.causal && !p.masked) softmax_block_causal_keep(r, mk.kinfo + bq * 32, qcmp, a2, mm2, kpw, t_p + bq * 16, l2);
          else softmax_block_generic(r, mk.kinfo + bq * 32, qi, a_i, m2, p.causal, p.masked, kpw, t_p + bq * 16, l);
        }
      };
      auto zero_until = [&](int from, int to) {                       // zero P for the skipped blocks in [from, to)
        for (int bq = from; bq < to; ++bq) {
          uint32_t z[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = 0u;
          tmem_st16(t_p + bq * 16, z);
        }
      };
      auto next_needed = [&](int after) -> int {                     // first needed block > after, or 4
        const uint32_t rest = need & ~((2u << after) - 1u);
        return rest ? __ffs(rest) - 1 : 4;
      };
      // The score buffer goes back to the issuer as soon as the LAST load of the pass has completed — before that
      // block's math, the row statistics and the wait for the P stores — so the next chunk's score MMA overlaps them.
      auto release_scores = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh.s_free[wg]);
      };
      int b0 = need ? __ffs(need) - 1 : 4;
      if (b0 < 4) tmem_ld32(t_lane + b0 * 32, ra);
      else release_scores();
      int done_to = 0;                                                // blocks < done_to are final
      while (b0 < 4) {
        const int b1 = next_needed(b0);
        tmem_ld_wait_dep(ra);
        if (b1 >= 4) release_scores();
        zero_until(done_to, b0);
        if (b1 < 4) tmem_ld32(t_lane + b1 * 32, rb);
        process(ra, b0);
        done_to = b0 + 1;
        if (b1 >= 4) break;
        const int b2 = next_needed(b1);
        tmem_ld_wait_dep(rb);
        if (b2 >= 4) release_scores();
        zero_until(done_to, b1);
        if (b2 < 4) tmem_ld32(t_lane + b2 * 32, ra);
        process(rb, b1);
        done_to = b1 + 1;
        b0 = b2;
      }
      zero_until(done_to, 4);
      if (tr2 && lane == 0) tr2[1] = clock64();
      if (warp == 4 && lane == 0) TC_TRACE(wk.k, 14);
      l += lo32(l2) + hi32(l2);
      // hand the row statistics to the epilogue warpgroup (slot ob was read by the epilogue two chunks back)
      mbar_wait(&sh.o_free[ob], (jo & 1) ^ 1);
      sh.row_l[ob][wg][row] = l;
      if (wg == 1) {
        sh.row_m2[ob][row] = m2; sh.row_off[ob][row] = lse_off;
        sh.row_tk[ob][row] = my_tk;
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      mbar_arrive(&sh.l_full[ob]);
      if (lane == 0) mbar_arrive(&sh.p_full[wg]);
      if (TC_TRACE_ON && lane == 0 && p.trace && blockIdx.x == 0 && wk.k < 120) atomicMax(reinterpret_cast<unsigned long long *>(p.trace) + wk.k * 16 + 3, static_cast<unsigned long long>(clock64()));
    }
  } else if (warp >= 8 && warp < 12) {
    // ================================ epilogue warpgroup ==============================================
    const int row = (warp & 3) * 32 + lane;
    const uint32_t t_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const float keep_mul = p.keep_bits ? __ldg(p.keep_scale) : 1.f;   // dropout multiplier 1 / (1 - rate), folded into 1 / l
    for (Walker wk(g0, g1, p.n_chunks); wk.valid(); wk.next()) {
      const uint32_t w = wk.k & 1, j = wk.k >> 1;
      const int u = wk.u, b = u / p.H, h = u - b * p.H;
      mbar_wait(&sh.l_full[w], j & 1);                      // acquire the softmax threads' row statistics
      mbar_wait(&sh.o_full[w], j & 1);
      tc_fence_after();
      if (row == 0) TC_TRACE(wk.k, 4);
      const float l = sh.row_l[w][0][row] + sh.row_l[w][1][row], m2 = sh.row_m2[w][row];
      const float il = l > 0.f ? keep_mul / l : 0.f;
      const float lse = l > 0.f ? (m2 + log2f(l)) * kLn2 + sh.row_off[w][row] : -3e9f;
      const int tk = sh.row_tk[w][row];
      // The shift m2 is the row's analytic self score, not the maximum over its visible keys: when every visible key scores
      // ~83 (natural log units) or more below it, the exponentials flush to zero (large-norm queries whose nearest visible
      // key is far away).  Such rows are queued and redone with the true maximum by attend_fwd_redo_kernel.
#if TC_FWD_REDO
      if (!(l >= kRedoBelow) || !(l < kRedoAbove)) {
        const int slot = atomicAdd(p.redo, 1);
        p.redo[2 + 2 * slot] = wk.u * p.n_chunks + wk.c;
        p.redo[3 + 2 * slot] = tk;
      }
#endif
      uint32_t r0[32], r1[32];
      tmem_ld32(t_row + TC_O_COL + w * 64, r0);
      tmem_ld32(t_row + TC_O_COL + w * 64 + 32, r1);
      tmem_ld_wait_dep(r0);
      tmem_ld_wait_dep(r1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh.o_free[w]);            // accumulator and statistics slot w may be reused
      const int round = tk / p.L, pos = tk - round * p.L;
      __nv_bfloat16 *dst = p.o + b * p.o_sb + h * p.o_sh + round * p.o_sr + pos * p.o_sp;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 v;
        v.x = pack_bf16(__uint_as_float(r0[8 * q4 + 0]) * il, __uint_as_float(r0[8 * q4 + 1]) * il);
        v.y = pack_bf16(__uint_as_float(r0[8 * q4 + 2]) * il, __uint_as_float(r0[8 * q4 + 3]) * il);
        v.z = pack_bf16(__uint_as_float(r0[8 * q4 + 4]) * il, __uint_as_float(r0[8 * q4 + 5]) * il);
        v.w = pack_bf16(__uint_as_float(r0[8 * q4 + 6]) * il, __uint_as_float(r0[8 * q4 + 7]) * il);
        *reinterpret_cast<uint4 *>(dst + q4 * 8) = v;
        v.x = pack_bf16(__uint_as_float(r1[8 * q4 + 0]) * il, __uint_as_float(r1[8 * q4 + 1]) * il);
        v.y = pack_bf16(__uint_as_float(r1[8 * q4 + 2]) * il, __uint_as_float(r1[8 * q4 + 3]) * il);
        v.z = pack_bf16(__uint_as_float(r1[8 * q4 + 4]) * il, __uint_as_float(r1[8 * q4 + 5]) * il);
        v.w = pack_bf16(__uint_as_float(r1[8 * q4 + 6]) * il, __uint_as_float(r1[8 * q4 + 7]) * il);
        *reinterpret_cast<uint4 *>(dst + 32 + q4 * 8) = v;
      }
      p.lse[static_cast<int64_t>(u) * p.N + tk] = lse;
      if (row == 0) TC_TRACE(wk.k, 5);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (TC_TRACE_ON && p.trace && tid == 0) p.trace[120 * 16 + blockIdx.x] = clock64() - t_cta_start;   // per-CTA duration (load balance)
  if (warp == 14) tmem_dealloc(tmem, 512);
}

// Exact redo of the rows the tensor-core kernel could not normalise (see the epilogue): one warp per queued row walks the
// row's window in the reference's slot order with fp32 arithmetic and the reference's subtractive masks (EA:229-231,
// 244, 148-159, 251-252, 254-265), true maximum first.  Rare by construction (the queue is usually empty).
__global__ void __launch_bounds__(256) attend_fwd_redo_kernel(const AttendFwdParams p) {
  const int lane = threadIdx.x & 31;
  const int n = p.redo[0];
  for (int e = blockIdx.x * 8 + (threadIdx.x >> 5); e < n; e += gridDim.x * 8) {
    const int item = p.redo[2 + 2 * e], tk = p.redo[3 + 2 * e];
    const int u = item / p.n_chunks, c = item - u * p.n_chunks, b = u / p.H, h = u - b * p.H;
    const int round = tk / p.L, pos = tk - round * p.L;
    const int32_t *stk = p.sticker + static_cast<int64_t>(u) * p.N;
    const __nv_bfloat16 *qv_u = p.qv + (static_cast<int64_t>(b) * p.L * p.H + h) * 128;
    const int64_t rstride = static_cast<int64_t>(p.H) * 128;
    const float2 q2 = unpack_bf16(*reinterpret_cast<const uint32_t *>(qv_u + pos * rstride + 2 * lane));
    const float qi = static_cast<float>(pos + 1);                      // q_info (EA:201)
    int qslot = 0;                                                     // my slot in the chunk (dropout row)
    for (int l4 = 0; l4 < 4; ++l4) {
      const unsigned hit = __ballot_sync(0xffffffffu, stk[c * TC_C + l4 * 32 + lane] == tk);
      if (hit) qslot = l4 * 32 + __ffs(hit) - 1;
    }
    float sc[8];                                                       // scores of keys j with j % 32 == lane
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      sc[jj] = -INFINITY;
      if (jj * 32 < p.nwin * TC_C) {
        const int cc = ((c + (jj >> 2) - p.nb) % p.n_chunks + p.n_chunks) % p.n_chunks;   // cyclic window (EA:137-141)
        const int ktk_l = stk[cc * TC_C + (jj & 3) * 32 + lane];
        for (int l = 0; l < 32; ++l) {
          const int kpos = __shfl_sync(0xffffffffu, ktk_l, l) % p.L;
          const float2 k2 = unpack_bf16(*reinterpret_cast<const uint32_t *>(qv_u + kpos * rstride + 2 * lane));
          float dot = q2.x * k2.x + q2.y * k2.y, ss = k2.x * k2.x + k2.y * k2.y;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            ss += __shfl_xor_sync(0xffffffffu, ss, o);
          }
          bool kvalid = true;
          if (p.masked) kvalid = p.mask[static_cast<int64_t>(b) * p.L + kpos] != 0;
          const float ki = static_cast<float>((kvalid ? kpos : -kpos) + 1);
          float s = dot / sqrtf(ss * (1.f / 64) + 1e-6f) * 0.125f;     // EA:229-231, 244
          if (p.causal && qi < ki) s -= 1e9f;                          // EA:150-152
          if (qi == ki) s -= 1e5f;                                     // EA:153-155
          if (p.masked && ki < 0.f) s -= 1e9f;                         // EA:156-159
          if (lane == l) sc[jj] = s;
        }
        mx = fmaxf(mx, sc[jj]);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float lsum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      sc[jj] = jj * 32 < p.nwin * TC_C ? __expf(sc[jj] - mx) : 0.f;
      lsum += sc[jj];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    const float inv = 1.f / lsum;
    float o0 = 0.f, o1 = 0.f;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      if (jj * 32 < p.nwin * TC_C) {
        const int cc = ((c + (jj >> 2) - p.nb) % p.n_chunks + p.n_chunks) % p.n_chunks;
        const int ktk_l = stk[cc * TC_C + (jj & 3) * 32 + lane];
        uint32_t keep = 0xffffffffu;
        if (p.keep_bits) keep = p.keep_bits[qslot * (p.nwin * TC_C / 32) + jj];
        for (int l = 0; l < 32; ++l) {
          const int kpos = __shfl_sync(0xffffffffu, ktk_l, l) % p.L;
          float pj = __shfl_sync(0xffffffffu, sc[jj], l) * inv;
          if (p.keep_bits) pj = (keep >> l) & 1u ? pj * __ldg(p.keep_scale) : 0.f;   // EA:254-262: after the softmax, lse untouched
          const float2 v2 = unpack_bf16(*reinterpret_cast<const uint32_t *>(qv_u + kpos * rstride + 64 + 2 * lane));
          o0 = fmaf(pj, v2.x, o0); o1 = fmaf(pj, v2.y, o1);
        }
      }
    }
    __nv_bfloat16 *dst = p.o + b * p.o_sb + h * p.o_sh + round * p.o_sr + pos * p.o_sp;
    *reinterpret_cast<uint32_t *>(dst + 2 * lane) = pack_bf16(o0, o1);
    if (lane == 0) p.lse[static_cast<int64_t>(u) * p.N + tk] = mx + __logf(lsum);
  }
}

bool attend_fwd_tc_uses_bounds() { return TC_FWD_BOUNDS != 0; }

int attend_fwd_tc_run(const AttendFwdParams &p, int BH, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(TC_NST) * TC_TILE_BYTES + 1024;   // + static TcShared
  const bool sorted = p.causal && !p.masked && p.nb == 1 && !p.keep_bits;
  LSH_OPT_IN_SMEM(attend_fwd_tc_kernel<true>);
  LSH_OPT_IN_SMEM(attend_fwd_tc_kernel<false>);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int total = BH * p.n_chunks;
  int grid = total < sms ? total : sms;
  if (const char *e = getenv("LSH_ATTN_MAX_CTAS")) {   // test hook: force many chunks per CTA on small problems
    const int m = atoi(e);
    if (m > 0 && m < grid) grid = m;
  }
  if (!p.redo) return set_error("attend_fwd_tc: redo queue missing from the workspace");
#if TC_FWD_TMA
  AttendFwdParams pt = p;
  const int64_t B = BH / p.H;
  if (int rc = make_row_gather_map(&pt.tm_k, p.qhat, static_cast<uint64_t>(BH) * p.L, 64, 128, 64)) return rc;
  if (int rc = make_row_gather_map(&pt.tm_v, p.qv, static_cast<uint64_t>(B) * p.L * p.H, 128, 256, 64)) return rc;
#else
  const AttendFwdParams &pt = p;
#endif
  if (cudaMemsetAsync(p.redo, 0, 8, stream) != cudaSuccess) return set_error("attend_fwd_tc: cudaMemsetAsync failed");
  if (sorted) attend_fwd_tc_kernel<true><<<grid, TC_THREADS, smem, stream>>>(pt, total);
  else attend_fwd_tc_kernel<false><<<grid, TC_THREADS, smem, stream>>>(pt, total);
  LSH_CHECK_LAUNCH("attend_fwd_tc_kernel");
  attend_fwd_redo_kernel<<<sms, 256, 0, stream>>>(p);
  LSH_CHECK_LAUNCH("attend_fwd_redo_kernel");
  return 0;
}

}  // namespace lsh
