// Non-causal flash attention forward for sm_100a, head_dim 128, bf16 operands, fp32 softmax.
//
// One CTA owns TWO 128-row Q tiles of one head and streams K/V tiles of 128 rows past them:
//
//   warps 0-3  : softmax warpgroup for Q tile 0 (thread = row; TMEM lane quadrant = warp % 4)   208 registers
//   warps 4-7  : softmax warpgroup for Q tile 1                                                 208 registers
//   warp  8    : TMA producer (Q once, then K_j / V_j into 2-stage rings, 128B-swizzled boxes)   88 registers
//   warp  9    : tcgen05.mma issuer + TMEM owner                                                 88 registers
//   warps 10-11: idle (they complete the third warpgroup: setmaxnreg is a warpgroup-wide instruction)
//
// What bounds it (ncu source view of the round-1 kernel, profiles/r02_attn_analysis.md): in the softmax loop every issued
// instruction carried ~1.1 cycles of `stall_wait` on top of its issue slot (FFMA2 / FADD2 / FMNMX3 / IMAD hold the port for two
// cycles), ~100 IMADs were register shuffles forced by the 168-register ceiling, and one SM sub-partition serves one warp of
// each softmax warpgroup — the round-1 loop needed ~2200 cycles per 128x128 tile and warp where the tensor pipe needs 1024.
// This version: `setmaxnreg` moves registers from the TMA / MMA / idle warpgroup (88) to the softmax warpgroups (208): no
// spills, 29 IMADs instead of 98; the share of exponentials computed by the FMA-pipe polynomial was re-swept with that
// (TF/s at L = 32760, H = 12, same box: 4/16 1391, 6/16 1378, 8/16 1334, 10/16 1227, 16/16 1028; all-MUFU 1315 on a slower
// box where 3/16 gave 1288) -> 5/16.  Same box, same run: round-1 kernel 1252 TF/s, this one 1411.
//
// TMEM (512 columns): S0 [0,128) | S1 [128,256) | O0 [256,384) | O1 [384,512); P_i (bf16, 64 columns)
// aliases the front of S_i and is consumed straight from TMEM by the P*V MMA (A operand in TMEM).
// The issue order  PV_i(j) ; QK_i(j+1)  per tile ping-pongs the two softmax warpgroups against the
// tensor pipe: while warpgroup 0 exponentiates S0(j+1), the tensor core runs PV1(j) and QK1(j+1).
// Online softmax uses the lazy-rescale rule (O is only rescaled when a row max grows by > 2^8).
//
// Work units are (head, pair of Q tiles).  The launch plan (plan_split, host side) runs `n_full` units whole and cuts
// each of the remaining ones into `split` slices of the K/V stream, so the last, partly filled wave of CTAs is spread
// over all SMs; sliced units leave (unnormalised O, row max, row sum) in a workspace and attn_merge_kernel combines
// them.  Without a workspace every unit runs whole.
// Replaces flash_attention(), reference wan_video_dit.py:116-147.
#include "common.cuh"
#include "../../include/svi_b200.h"

namespace svi {
namespace attn {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int HD = 128;
constexpr int KV_STAGES = 2;
constexpr int HALF_BYTES = 128 * 64 * 2;   // one 128-row x 64-col swizzled box (16 KB)
constexpr int TILE_BYTES = 2 * HALF_BYTES;  // 128 x 128 bf16 (32 KB)
constexpr int NUM_THREADS = 384;   // 12 warps = 3 warpgroups: setmaxnreg is a warpgroup-wide instruction
#ifndef SVI_ATTN_POLY16
#define SVI_ATTN_POLY16 5            // exponentials on the FMA pipes: this many of every 16 element pairs
#endif
#ifndef SVI_ATTN_COLSPLIT
#define SVI_ATTN_COLSPLIT 0         // 1: both softmax warpgroups work on the same Q tile (64 key columns each)
#endif
#ifndef SVI_ATTN_OTHER_REGS
#define SVI_ATTN_OTHER_REGS 88
#endif
#ifndef SVI_ATTN_SOFTMAX_REGS
#define SVI_ATTN_SOFTMAX_REGS 208
#endif
// setmaxnreg only redistributes the registers the CTA was launched with (384 threads x 168): a larger request than the
// pool holds blocks forever
static_assert(2 * SVI_ATTN_SOFTMAX_REGS + SVI_ATTN_OTHER_REGS <= 3 * 168, "register split exceeds the CTA's allocation");
static_assert(SVI_ATTN_SOFTMAX_REGS % 8 == 0 && SVI_ATTN_OTHER_REGS % 8 == 0, "setmaxnreg takes multiples of 8");
constexpr int TMEM_COLS = 512;
constexpr int XCH_BYTES = 2 * 2 * 2 * 128 * 4;   // row-max exchange between the two halves of a row: [parity][tile][half][row]
constexpr int SMEM_BYTES = (2 + 2 * KV_STAGES) * TILE_BYTES + 1024 + 256 + XCH_BYTES;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 of two values on the FMA/ALU pipes (Cody-Waite range reduction + degree-3 minimax polynomial on [-0.5, 0.5],
// max relative error 7.5e-5 — invisible after the bf16 rounding of P).  MUFU.EX2 issues at 16 lanes/clk/SM, i.e. the
// 2 x 128 x 128 exponentials of one K/V step cost as many cycles as its four MMAs; moving ~3/8 of them here balances
// the XU, FMA and ALU pipes.  Uses the sm_100 packed-pair fp32 instructions (FFMA2 / FADD2).
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));   // 1.5 * 2^23: round(x) lands in the mantissa
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);             // x - round(x) in [-0.5, 0.5]
  float2 q = __ffma2_rn(f, make_float2(0.055170949548482895f, 0.055170949548482895f),
                        make_float2(0.2426096349954605f, 0.2426096349954605f));
  q = __ffma2_rn(q, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  q = __ffma2_rn(q, f, make_float2(0.9999281764030457f, 0.9999281764030457f));
  q.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));   // * 2^round(x)
  q.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return q;
}

struct Params {
  __nv_bfloat16* O;
  long long ldo;
  int Lq, Lk;
  float scale_log2;  // softmax scale * log2(e)
  int accumulate;
  // sequence-parallel K/V stream (sp_exchange.cu); kv_flags == nullptr: plain attention
  const uint32_t* kv_flags;  // [n_chunks] arrival flag of each source rank's rows, written by that rank's push
  uint32_t kv_epoch;         // value a flag holds once the rows of this launch have landed
  int kv_chunk_rows;         // rows owned by each rank
  int kv_self_chunk;         // this rank's chunk: produced locally, never waited for
  int kv_first_tile;         // KV tile the stream starts on (first tile fully inside the local chunk)
  // launch plan: units [0, n_full) whole; unit n_full + t/split gets slice t%split of the K/V stream
  int n_qpairs;              // Q-tile pairs per head
  int n_full;
  int split;
  float* ws_o;               // [slices][256 rows][128] unnormalised partial O
  float2* ws_ml;             // [slices][256 rows] (row max in scaled log2 units, row sum)
  // optional per-row RMS scale of Q (full-width RMSNorm folded into the softmax scale): row r of Q is used as
  // Q[r] * rsqrt(sum_{j < q_ss_parts} q_sumsq[r * q_ss_ld + j] * q_inv_d + q_eps); nullptr: none
  const float* q_sumsq;
  int q_ss_ld, q_ss_parts;
  float q_inv_d, q_eps;
};

__device__ __forceinline__ float q_row_sumsq(const Params& p, int qrow) {   // partials added in index order: reproducible
  const float* sp = p.q_sumsq + (long long)qrow * p.q_ss_ld;
  float s = __ldg(sp);
  for (int j = 1; j < p.q_ss_parts; ++j) s += __ldg(sp + j);
  return s;
}

// KV tile visited at iteration j: the stream starts on the rank's own rows and wraps around
__device__ __forceinline__ int kv_tile_at(int j, int first_tile, int n_kv) {
  const int t = j + first_tile;
  return t >= n_kv ? t - n_kv : t;
}

// Producer side of the K/V exchange: block until the rows of `chunk` have been pushed into this GPU's buffer.
__device__ __forceinline__ void wait_kv_chunk(const uint32_t* flags, int chunk, uint32_t epoch) {
  const uint32_t* f = flags + chunk;
  uint32_t v;
  uint32_t spins = 0;
  while (true) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v == epoch) break;
    __nanosleep(64);
    if (++spins > (1u << 25)) {   // > 2 s: a peer died or the protocol is broken -> visible error, not a hang
      printf("svi: K/V exchange timeout block(%d,%d) chunk %d flag %u want %u\n", blockIdx.x, blockIdx.y, chunk, v, epoch);
      __trap();
    }
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");   // later TMA (async proxy) reads see the pushed rows
}

// Register budget: the kernel starts with 65536 / 384 = 168 registers per thread; once the roles are fixed the TMA / MMA /
// idle warpgroup drops to SVI_ATTN_OTHER_REGS and the two softmax warpgroups grow to SVI_ATTN_SOFTMAX_REGS
// (256*208 + 128*88 = 64512 = the launch allocation: setmaxnreg can only redistribute what the CTA already owns).
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SVI_ATTN_SOFTMAX_REGS));
}
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SVI_ATTN_OTHER_REGS)); }

// ---- shared-memory map (32-bit shared addresses throughout: the role code keeps no 64-bit generic pointers alive) ----
constexpr uint32_t OFF_Q = 0;                                   // [2][TILE_BYTES]
constexpr uint32_t OFF_K = 2 * TILE_BYTES;                      // [KV_STAGES][TILE_BYTES]
constexpr uint32_t OFF_V = OFF_K + KV_STAGES * TILE_BYTES;      // [KV_STAGES][TILE_BYTES]
constexpr uint32_t OFF_BAR = OFF_V + KV_STAGES * TILE_BYTES;
constexpr uint32_t OFF_XCH = OFF_BAR + 256;
enum : uint32_t {
  Q_FULL = 0,    // [2]
  K_FULL = 2,    // [2]
  K_EMPTY = 4,   // [2]
  V_FULL = 6,    // [2]
  V_EMPTY = 8,   // [2]
  S_FULL = 10,   // [2]  MMA -> softmax_i : S_i(j) ready (and PV_i(j-1) retired)
  P_READY = 12,  // [2][2] softmax_i -> MMA : half h (64 keys) of P_i(j) in TMEM, O_i rescaled; the P*V of the first
                 //        half overlaps the exponentials of the second half
  O_FULL = 16,   // [2]  MMA -> softmax_i : O_i final
  NUM_BARS = 18
};

struct Unit {      // what this CTA computes (decoded per role: nothing of it stays live across the role dispatch)
  int head, q_row0, j_begin, n_kv, n_kv_total, slice_slot;
};
__device__ __forceinline__ Unit decode_unit(const Params& p) {
  Unit u;
  u.n_kv_total = (p.Lk + BKV - 1) / BKV;
  int unit = blockIdx.x;
  u.j_begin = 0;
  u.n_kv = u.n_kv_total;                            // K/V tiles THIS CTA streams, starting at j_begin
  u.slice_slot = (int)blockIdx.x - p.n_full;        // >= 0: this CTA computes one slice of a unit
  if (u.slice_slot >= 0) {
    unit = p.n_full + u.slice_slot / p.split;
    const int sl = u.slice_slot % p.split;
    u.j_begin = (int)((long long)u.n_kv_total * sl / p.split);
    u.n_kv = (int)((long long)u.n_kv_total * (sl + 1) / p.split) - u.j_begin;
  }
  u.head = unit / p.n_qpairs;
  u.q_row0 = (unit % p.n_qpairs) * (2 * BQ);
  return u;
}

// ------------------------------------ TMA producer (one lane) ------------------------------------
__device__ __forceinline__ void tma_role(const CUtensorMap* tmap_q, const CUtensorMap* tmap_k, const CUtensorMap* tmap_v,
                                         const Params& p, uint32_t sbase) {
  const Unit u = decode_unit(p);
  const uint32_t bars = sbase + OFF_BAR;
  const int col0 = u.head * HD;
  auto load_tile = [&](uint32_t dst, const CUtensorMap* m, uint32_t bar, int row) {
    mbar_expect_tx_a(bar, TILE_BYTES);
    tma_load_2d_a(dst, m, bar, col0, row);
    tma_load_2d_a(dst + HALF_BYTES, m, bar, col0 + 64, row);
  };
  load_tile(sbase + OFF_Q, tmap_q, bars + 8 * Q_FULL, u.q_row0);
  int landed = p.kv_self_chunk;   // most recent remote chunk known to be present (chunks are visited in runs)
  for (int j = 0; j < u.n_kv; ++j) {
    const int s = j % KV_STAGES;
    const uint32_t ph = (j / KV_STAGES) & 1;
    const int row = kv_tile_at(u.j_begin + j, p.kv_first_tile, u.n_kv_total) * BKV;
    if (p.kv_flags) {
      const int c0 = row / p.kv_chunk_rows;
      const int c1 = (min(row + BKV, p.Lk) - 1) / p.kv_chunk_rows;
      for (int c = c0; c <= c1; ++c) {
        if (c == p.kv_self_chunk || c == landed) continue;
        wait_kv_chunk(p.kv_flags, c, p.kv_epoch);
        landed = c;
      }
    }
    mbar_wait_a(bars + 8 * (K_EMPTY + s), ph ^ 1);
    load_tile(sbase + OFF_K + s * TILE_BYTES, tmap_k, bars + 8 * (K_FULL + s), row);
    if (j == 0) load_tile(sbase + OFF_Q + TILE_BYTES, tmap_q, bars + 8 * (Q_FULL + 1), u.q_row0 + BQ);
    mbar_wait_a(bars + 8 * (V_EMPTY + s), ph ^ 1);
    load_tile(sbase + OFF_V + s * TILE_BYTES, tmap_v, bars + 8 * (V_FULL + s), row);
  }
}

// ------------------------------------ MMA issuer (whole warp, uniform issue) ----------------------
__device__ __forceinline__ void mma_role(const Params& p, uint32_t sbase, uint32_t tmem_base) {
  const Unit u = decode_unit(p);
  const int n_kv = u.n_kv;
  const uint32_t bars = sbase + OFF_BAR;
  constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major (d contiguous)
  constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, HD, 0, 1);   // P K-major (TMEM), V MN-major
  constexpr uint32_t hi_kmaj = smem_desc_hi(1024, 2);            // SBO 1024 B (8 rows x 128 B), 128B swizzle
  const uint32_t q_lo = smem_desc_lo(sbase + OFF_Q, 16);
  const uint32_t k_lo = smem_desc_lo(sbase + OFF_K, 16);
  const uint32_t v_lo = smem_desc_lo(sbase + OFF_V, HALF_BYTES);  // MN-major: LBO = distance of the 64-col atoms
  // whole warp executes (uniform operands -> uniform registers); elect.sync inside the wrappers picks the issuing lane
  auto issue_qk = [&](int i, int ks) {   // 8 K-steps = 2 swizzle boxes x 4 (batched issue)
    const uint32_t a0 = q_lo + ((i * TILE_BYTES) >> 4), b0 = k_lo + ((ks * TILE_BYTES) >> 4);
    tc_mma_ss_k4(tmem_base + i * 128, a0, hi_kmaj, b0, hi_kmaj, idesc_qk, 0);
    tc_mma_ss_k4(tmem_base + i * 128, a0 + (HALF_BYTES >> 4), hi_kmaj, b0 + (HALF_BYTES >> 4), hi_kmaj, idesc_qk, 1);
  };
  auto issue_pv_half = [&](int i, int vs, int h, uint32_t accumulate) {   // 4 K-steps = 64 K/V rows of half h
    tc_mma_ts_k4(tmem_base + 256 + i * 128, tmem_base + i * 128 + h * 32, v_lo + ((vs * TILE_BYTES + h * 4 * 2048) >> 4),
                 hi_kmaj, idesc_pv, accumulate);
  };

  // prologue: S_i(0) = Q_i K_0^T
  mbar_wait_a(bars + 8 * K_FULL, 0);
  for (int i = 0; i < 2; ++i) {
    mbar_wait_a(bars + 8 * (Q_FULL + i), 0);
    tc_fence_after();
    issue_qk(i, 0);
    tc_commit_a(bars + 8 * (S_FULL + i));
  }
  tc_commit_a(bars + 8 * K_EMPTY);

  for (int j = 0; j < n_kv; ++j) {
    const int vs = j % KV_STAGES;
    const int ks = (j + 1) % KV_STAGES;
    const bool has_next = (j + 1) < n_kv;
    mbar_wait_a(bars + 8 * (V_FULL + vs), (j / KV_STAGES) & 1);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_wait_a(bars + 8 * (P_READY + i * 2 + 0), j & 1);
      tc_fence_after();
      issue_pv_half(i, vs, 0, j > 0);
      if (has_next && i == 0) mbar_wait_a(bars + 8 * (K_FULL + ks), ((j + 1) / KV_STAGES) & 1);
      mbar_wait_a(bars + 8 * (P_READY + i * 2 + 1), j & 1);
      tc_fence_after();
      issue_pv_half(i, vs, 1, 1);
      if (has_next) {
        issue_qk(i, ks);
        tc_commit_a(bars + 8 * (S_FULL + i));
      } else {
        tc_commit_a(bars + 8 * (O_FULL + i));
      }
    }
    tc_commit_a(bars + 8 * (V_EMPTY + vs));
    if (has_next) tc_commit_a(bars + 8 * (K_EMPTY + ks));
  }
}

#if SVI_ATTN_COLSPLIT
// ------------------------------------ softmax warpgroups -------------------------------------------
// Column split: warpgroup h (warps 4h..4h+3) owns key columns [64h, 64h+64) of BOTH Q tiles' S; thread = one row x 64 keys.
// Warps w and w+4 hold the same 32 rows (same TMEM lane quadrant, same SM sub-partition) and work on the same tile at the
// same time, so the S -> softmax -> P leg of a tile's dependency chain takes half as long as with one warpgroup per tile
// (the tensor pipe was idle ~40 % of the time waiting for that leg: profiles/r02_attn_analysis.md).  The two halves of a
// row exchange their maxima through shared memory behind a 64-thread named barrier (one per tile and quadrant); that
// barrier also orders "warpgroup 0 has loaded S columns 32..63" before "warpgroup 1 overwrites them with its half of P".
// Each warpgroup hands its 64-key half of P to the MMA warp on its own p_ready barrier, keeps its half of the row sum and
// rescales / normalises / stores its 64 output columns.
__device__ __forceinline__ void pair_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__device__ __forceinline__ void softmax_role(const Params& p, uint32_t sbase, uint32_t tmem_base, int warp, int lane) {
  const Unit u = decode_unit(p);
  const int n_kv = u.n_kv;
  const uint32_t bars = sbase + OFF_BAR;
  const int h = warp >> 2;    // which 64-key half of every S tile
  const int quad = warp & 3;  // TMEM lane quadrant = 32 rows of both Q tiles
  const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
  const int rl = quad * 32 + lane;                                  // row inside a Q tile
  float* xch = reinterpret_cast<float*>(__cvta_shared_to_generic(sbase + OFF_XCH));     // [2 parity][2 tile][2 half][128 rows]
  float c[2], m_cur[2], l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    c[i] = p.scale_log2;
    if (p.q_sumsq) {          // this thread's Q rows carry their RMSNorm factor in the softmax scale
      const int qrow = min(u.q_row0 + i * BQ + rl, p.Lq - 1);
      c[i] *= rsqrtf(q_row_sumsq(p, qrow) * p.q_inv_d + p.q_eps);
    }
    m_cur[i] = -INFINITY;     // running row max (scaled, log2 domain); reference point of P and O; identical in both halves
    l[i] = 0.f;               // running sum of THIS half's P
  }

  for (int j = 0; j < n_kv; ++j) {
    const int limit = p.Lk - kv_tile_at(u.j_begin + j, p.kv_first_tile, u.n_kv_total) * BKV - h * 64;  // valid columns of my half
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t tS = tmem_base + i * 128 + lane_sel;
      const uint32_t tO = tmem_base + 256 + i * 128 + lane_sel;
      mbar_wait_a(bars + 8 * (S_FULL + i), j & 1);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld32(tS + h * 64, sr[0]);
      tmem_ld32(tS + h * 64 + 32, sr[1]);
      tmem_ld_wait();
      if (limit < 64) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (cc * 32 + e >= limit) sr[cc][e] = 0xff800000u;  // -inf: masked key column
      }
      float m8[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) m8[q] = fmaxf(__uint_as_float(sr[0][q]), __uint_as_float(sr[0][q + 8]));
#pragma unroll
      for (int e = 16; e < 32; ++e) m8[e & 7] = fmaxf(m8[e & 7], __uint_as_float(sr[0][e]));
#pragma unroll
      for (int e = 0; e < 32; ++e) m8[e & 7] = fmaxf(m8[e & 7], __uint_as_float(sr[1][e]));
      const float mloc = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      // exchange with the other half of the row (warp ^ 4): double-buffered by step parity
      float* xr = xch + (((j & 1) * 2 + i) * 2) * 128;
      xr[h * 128 + rl] = mloc;
      pair_bar(1 + i * 4 + quad);
      const float mx = fmaxf(mloc, xr[(h ^ 1) * 128 + rl]) * c[i];
      // lazy rescale (warp-uniform decision; both halves of a row see the same mx, hence take the same decision)
      const bool need = (j > 0) && (mx > m_cur[i] + RESCALE_THRESHOLD);
      if (j == 0) {
        m_cur[i] = mx;
      } else if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_cur[i], mx);
        const float alpha = ex2(m_cur[i] - m_new);
        l[i] *= alpha;
        m_cur[i] = m_new;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {          // my 64 output columns
          uint32_t r[32];
          tmem_ld32(tO + h * 64 + cc * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
          tmem_st32(tO + h * 64 + cc * 32, r);
        }
        tmem_st_wait();
        pair_bar(1 + i * 4 + quad);               // the P*V of either half accumulates into ALL 128 columns of O
      }
      // P = exp2(S*c - m) (masked columns: exp2(-inf) = 0), bf16 pack into TMEM columns [32h, 32h+32) of the tile
      float2 l4[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float2 c2 = make_float2(c[i], c[i]), nm2 = make_float2(-m_cur[i], -m_cur[i]);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const int pr = e >> 1;   // pair index 0..15 inside the chunk
          float2 x = __ffma2_rn(make_float2(__uint_as_float(sr[cc][e]), __uint_as_float(sr[cc][e + 1])), c2, nm2);
          float2 pv;
          if (pr < SVI_ATTN_POLY16) {   // FMA/ALU-pipe exponential
            pv = exp2_poly2(x);
          } else {                      // MUFU exponential
            pv.x = ex2(x.x);
            pv.y = ex2(x.y);
          }
          l4[pr & 3] = __fadd2_rn(l4[pr & 3], pv);
          pk[pr] = pack_bf16x2(pv.x, pv.y);
        }
        tmem_st16(tS + h * 32 + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bars + 8 * (P_READY + i * 2 + h));
      {
        const float2 a = __fadd2_rn(__fadd2_rn(l4[0], l4[1]), __fadd2_rn(l4[2], l4[3]));
        l[i] += a.x + a.y;
      }
    }
  }

  // epilogue: O / l -> bf16 -> global (my 64 columns of both tiles)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t tO = tmem_base + 256 + i * 128 + lane_sel;
    mbar_wait_a(bars + 8 * (O_FULL + i), 0);
    tc_fence_after();
    float* xr = xch + ((n_kv & 1) * 2 + i) * 2 * 128;     // parity after the last step: not in use by a straggling half
    xr[h * 128 + rl] = l[i];
    pair_bar(1 + i * 4 + quad);
    const float lsum = l[i] + xr[(h ^ 1) * 128 + rl];
    const int row = u.q_row0 + i * BQ + rl;
    if (u.slice_slot >= 0 && p.split > 1) {
      // one slice of the K/V stream: leave (O, m, l) for attn_merge_kernel
      const long long prow = (long long)u.slice_slot * (2 * BQ) + i * BQ + rl;
      float4* dst = reinterpret_cast<float4*>(p.ws_o + prow * HD + h * 64);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        tmem_ld32(tO + h * 64 + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g)
          dst[cc * 8 + g] = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]),
                                        __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3]));
      }
      if (h == 0) p.ws_ml[prow] = make_float2(m_cur[i], lsum);
    } else {
      const float inv_l = 1.0f / lsum;
      __nv_bfloat16* orow = p.O + (long long)row * p.ldo + u.head * HD + h * 64;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        tmem_ld32(tO + h * 64 + cc * 32, r);
        tmem_ld_wait();
        if (row < p.Lq) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[g * 8 + e]) * inv_l;
            uint4* dst = reinterpret_cast<uint4*>(orow + cc * 32 + g * 8);
            if (p.accumulate) {
              const uint4 old = *dst;
              const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(ob[e]);
                v[2 * e] += f.x;
                v[2 * e + 1] += f.y;
              }
            }
            uint4 pk;
            pk.x = pack_bf16x2(v[0], v[1]);
            pk.y = pack_bf16x2(v[2], v[3]);
            pk.z = pack_bf16x2(v[4], v[5]);
            pk.w = pack_bf16x2(v[6], v[7]);
            *dst = pk;
          }
        }
      }
    }
  }
}

#else
// ------------------------------------ softmax warpgroups (thread = one row of Q tile i) -----------
__device__ __forceinline__ void softmax_role(const Params& p, uint32_t sbase, uint32_t tmem_base, int warp, int lane) {
  const Unit u = decode_unit(p);
  const int n_kv = u.n_kv;
  const uint32_t bars = sbase + OFF_BAR;
  const int i = warp >> 2;    // which Q tile
  const int quad = warp & 3;  // TMEM lane quadrant
  const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
  const uint32_t tS = tmem_base + i * 128 + lane_sel;
  const uint32_t tO = tmem_base + 256 + i * 128 + lane_sel;
  float c = p.scale_log2;
  if (p.q_sumsq) {          // this thread's Q row carries its RMSNorm factor in the softmax scale
    const int qrow = min(u.q_row0 + i * BQ + quad * 32 + lane, p.Lq - 1);
    c *= rsqrtf(q_row_sumsq(p, qrow) * p.q_inv_d + p.q_eps);
  }
  float m_cur = -INFINITY;  // running row max (scaled, log2 domain); reference point of P and O
  float l = 0.f;            // running row sum of P

  for (int j = 0; j < n_kv; ++j) {
    mbar_wait_a(bars + 8 * (S_FULL + i), j & 1);
    tc_fence_after();
    const int limit = p.Lk - kv_tile_at(u.j_begin + j, p.kv_first_tile, u.n_kv_total) * BKV;  // valid key columns (>=128: all)
    // single pass: the whole 128-wide S row of this thread lives in registers (4 TMEM loads in flight, one wait)
    uint32_t sr[4][32];
    tmem_ld32(tS + 0, sr[0]);
    tmem_ld32(tS + 32, sr[1]);
    tmem_ld32(tS + 64, sr[2]);
    tmem_ld32(tS + 96, sr[3]);
    tmem_ld_wait();
    if (limit < BKV) {
#pragma unroll
      for (int cc = 0; cc < 4; ++cc)
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (cc * 32 + e >= limit) sr[cc][e] = 0xff800000u;  // -inf: masked key column
    }
    // row max: 8 independent chains (the 128-long serial fmax chain was the critical path)
    float m8[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) m8[q] = fmaxf(__uint_as_float(sr[0][q]), __uint_as_float(sr[0][q + 8]));
#pragma unroll
    for (int e = 16; e < 32; ++e) m8[e & 7] = fmaxf(m8[e & 7], __uint_as_float(sr[0][e]));
#pragma unroll
    for (int cc = 1; cc < 4; ++cc) {
#pragma unroll
      for (int e = 0; e < 32; ++e) m8[e & 7] = fmaxf(m8[e & 7], __uint_as_float(sr[cc][e]));
    }
    float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
    mx *= c;
    // lazy rescale (warp-uniform decision because tcgen05.ld/st are warp-collective)
    const bool need = (j > 0) && (mx > m_cur + RESCALE_THRESHOLD);
    if (j == 0) {
      m_cur = mx;
    } else if (__any_sync(0xffffffffu, need)) {
      const float m_new = fmaxf(m_cur, mx);
      const float alpha = ex2(m_cur - m_new);
      l *= alpha;
      m_cur = m_new;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld32(tO + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * alpha);
        tmem_st32(tO + cc * 32, r);
      }
    }
    // P = exp2(S*c - m) (masked columns: exp2(-inf) = 0), 4 independent row-sum chains, bf16 pack into TMEM
    float2 l4[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_cur, -m_cur);
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const int pr = e >> 1;   // pair index 0..15 inside the chunk
        float2 x = __ffma2_rn(make_float2(__uint_as_float(sr[cc][e]), __uint_as_float(sr[cc][e + 1])), c2, nm2);
        float2 pv;
        if (pr < SVI_ATTN_POLY16) {   // FMA/ALU-pipe exponential
          pv = exp2_poly2(x);
        } else {                      // MUFU exponential
          pv.x = ex2(x.x);
          pv.y = ex2(x.y);
        }
        l4[pr & 3] = __fadd2_rn(l4[pr & 3], pv);
        pk[pr] = pack_bf16x2(pv.x, pv.y);
      }
      tmem_st16(tS + cc * 16, pk);
      if (cc & 1) {   // a 64-key half of P is complete: hand it to the MMA warp now
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_a(bars + 8 * (P_READY + i * 2 + (cc >> 1)));
      }
    }
    {
      const float2 a = __fadd2_rn(__fadd2_rn(l4[0], l4[1]), __fadd2_rn(l4[2], l4[3]));
      l += a.x + a.y;
    }
  }

  // epilogue: O / l -> bf16 -> global
  mbar_wait_a(bars + 8 * (O_FULL + i), 0);
  tc_fence_after();
  const int row = u.q_row0 + i * BQ + quad * 32 + lane;
  if (u.slice_slot >= 0 && p.split > 1) {
    // one slice of the K/V stream: leave (O, m, l) for attn_merge_kernel
    const long long prow = (long long)u.slice_slot * (2 * BQ) + i * BQ + quad * 32 + lane;
    float4* dst = reinterpret_cast<float4*>(p.ws_o + prow * HD);
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t r[32];
      tmem_ld32(tO + cc * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 8; ++g)
        dst[cc * 8 + g] = make_float4(__uint_as_float(r[g * 4]), __uint_as_float(r[g * 4 + 1]),
                                      __uint_as_float(r[g * 4 + 2]), __uint_as_float(r[g * 4 + 3]));
    }
    p.ws_ml[prow] = make_float2(m_cur, l);
  } else {
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = p.O + (long long)row * p.ldo + u.head * HD;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t r[32];
      tmem_ld32(tO + cc * 32, r);
      tmem_ld_wait();
      if (row < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[g * 8 + e]) * inv_l;
          uint4* dst = reinterpret_cast<uint4*>(orow + cc * 32 + g * 8);
          if (p.accumulate) {
            const uint4 old = *dst;
            const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(ob[e]);
              v[2 * e] += f.x;
              v[2 * e + 1] += f.y;
            }
          }
          uint4 pk;
          pk.x = pack_bf16x2(v[0], v[1]);
          pk.y = pack_bf16x2(v[2], v[3]);
          pk.z = pack_bf16x2(v[4], v[5]);
          pk.w = pack_bf16x2(v[6], v[7]);
          *dst = pk;
        }
      }
    }
  }
}

#endif

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t tmem_slot = sbase + OFF_BAR + 8 * NUM_BARS;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 9) {
    if (lane == 0) {
      const uint32_t bars = sbase + OFF_BAR;
      for (uint32_t b = 0; b < NUM_BARS; ++b)
        mbar_init_a(bars + 8 * b, (b >= P_READY && b < O_FULL) ? 4u : 1u);   // P_READY: one arrive per softmax warp
      fence_mbar_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp >= 8) {
    setmaxnreg_dec();
    if (warp == 8) {
      if (lane == 0) tma_role(&tmap_q, &tmap_k, &tmap_v, p, sbase);
    } else if (warp == 9) {
      mma_role(p, sbase, tmem_base);
    }
  } else {
    setmaxnreg_inc();
    softmax_role(p, sbase, tmem_base, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace attn
}  // namespace svi

namespace svi {
namespace attn {

constexpr size_t SLICE_BYTES = (size_t)2 * BQ * HD * 4 + (size_t)2 * BQ * 8;   // partial O + (m, l) of one slice
constexpr int MAX_SPLIT = 8;
constexpr int MIN_SLICE_TILES = 8;   // a slice shorter than this is dominated by its prologue / epilogue

// Combines the slices of the sliced units: one warp per Q row (lane = 4 output columns), 8 rows per block.
__global__ void __launch_bounds__(256)
attn_merge_kernel(const float* __restrict__ ws_o, const float2* __restrict__ ws_ml, int split, int n_full, int n_qpairs,
                  int Lq, __nv_bfloat16* __restrict__ O, long long ldo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x / (2 * BQ / 8);                       // sliced unit
  const int r = (blockIdx.x % (2 * BQ / 8)) * 8 + warp;          // row inside the unit's 256
  const int unit = n_full + t;
  const int head = unit / n_qpairs;
  const int row = (unit % n_qpairs) * (2 * BQ) + r;
  if (row >= Lq) return;
  float m = -INFINITY;
  for (int s = 0; s < split; ++s) m = fmaxf(m, ws_ml[((long long)t * split + s) * (2 * BQ) + r].x);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float l = 0.f;
  for (int s = 0; s < split; ++s) {
    const long long prow = ((long long)t * split + s) * (2 * BQ) + r;
    const float2 ml = ws_ml[prow];
    const float w = ex2(ml.x - m);
    const float4 o = __ldcs(reinterpret_cast<const float4*>(ws_o + prow * HD) + lane);
    acc.x += o.x * w; acc.y += o.y * w; acc.z += o.z * w; acc.w += o.w * w;
    l += ml.y * w;
  }
  const float inv = 1.0f / l;
  uint2 pk;
  pk.x = pack_bf16x2(acc.x * inv, acc.y * inv);
  pk.y = pack_bf16x2(acc.z * inv, acc.w * inv);
  *reinterpret_cast<uint2*>(O + (long long)row * ldo + head * HD + lane * 4) = pk;
}

// Launch plan: `units` equal CTAs on `sms` SMs cost ceil(units/sms) rounds.  Running r rounds of whole units and cutting
// the remaining t units into S slices costs r + ceil(t*S/sms)/S rounds (+ a little for the merge).  Returns the cheapest
// (n_full, split) that fits the workspace; (units, 1) when slicing does not pay.
static void plan_split(int units, int n_kv, int sms, size_t ws_bytes, int* n_full, int* split) {
  *n_full = units;
  *split = 1;
  if (ws_bytes < SLICE_BYTES || units <= 0) return;
  const int rounds = units / sms;
  double best = (double)((units + sms - 1) / sms);
  const double need = best * 0.985;   // must save at least 1.5 %
  for (int r = rounds; r >= 0 && r >= rounds - 1; --r) {
    const int t = units - r * sms;
    if (t <= 0) continue;
    for (int S = 2; S <= MAX_SPLIT; ++S) {
      if (n_kv / S < MIN_SLICE_TILES) break;
      const long long slices = (long long)t * S;
      if ((size_t)slices * SLICE_BYTES > ws_bytes) break;
      const double cost = r + (double)((slices + sms - 1) / sms) / S + 0.04 * (double)t / sms + 0.01 * S;
      if (cost < best && cost < need) {
        best = cost;
        *n_full = units - t;
        *split = S;
      }
    }
  }
}

struct QScale {
  const float* sumsq = nullptr;
  int ld = 0;
  int parts = 1;
  int dim = 1;
  float eps = 0.f;
};

static int launch_attn(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                       int64_t ldo, int32_t Lq, int32_t Lk, int32_t num_heads, float scale, int32_t accumulate,
                       const uint32_t* kv_flags, uint32_t kv_epoch, int kv_chunk_rows, int kv_self_chunk, void* workspace,
                       size_t workspace_bytes, void* stream, const char* who, QScale qs = QScale()) {
  SVI_REQUIRE(Q && K && V && O, "%s: null pointer", who);
  SVI_REQUIRE(Lq > 0 && Lk > 0 && num_heads > 0, "%s: Lq, Lk, num_heads must be positive", who);
  const int64_t width = (int64_t)num_heads * HD;
  SVI_REQUIRE(ldq >= width && ldk >= width && ldv >= width && ldo >= width,
              "%s: leading dimensions must be >= num_heads*128", who);
  SVI_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
              "%s: leading dimensions must be multiples of 8 elements", who);
  SVI_REQUIRE(((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(K) |
                reinterpret_cast<uintptr_t>(V) | reinterpret_cast<uintptr_t>(O)) & 15) == 0,
              "%s: pointers must be 16-byte aligned", who);
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_2d(&tq, Q, 2, (uint64_t)width, (uint64_t)Lq, (uint64_t)ldq * 2, 64, BQ);
  if (rc) return rc;
  rc = make_tmap_2d(&tk, K, 2, (uint64_t)width, (uint64_t)Lk, (uint64_t)ldk * 2, 64, BKV);
  if (rc) return rc;
  rc = make_tmap_2d(&tv, V, 2, (uint64_t)width, (uint64_t)Lk, (uint64_t)ldv * 2, 64, BKV);
  if (rc) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          SMEM_BYTES);
    if (ce != cudaSuccess) {
      set_last_error("%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(ce));
      return SVI_ERR_LAUNCH;
    }
    attr_set = true;
  }
  Params p;
  p.O = reinterpret_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.Lq = Lq;
  p.Lk = Lk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.accumulate = accumulate;
  p.kv_flags = kv_flags;
  p.kv_epoch = kv_epoch;
  p.kv_chunk_rows = kv_chunk_rows;
  p.kv_self_chunk = kv_self_chunk;
  p.q_sumsq = qs.sumsq;
  p.q_ss_ld = qs.ld;
  p.q_ss_parts = qs.parts;
  p.q_inv_d = 1.0f / (float)qs.dim;
  p.q_eps = qs.eps;
  p.kv_first_tile = 0;
  if (kv_flags) {
    const int n_kv = (Lk + BKV - 1) / BKV;
    const int first = (int)(((int64_t)kv_self_chunk * kv_chunk_rows + BKV - 1) / BKV);
    p.kv_first_tile = first >= n_kv ? 0 : first;
  }
  p.n_qpairs = (Lq + 2 * BQ - 1) / (2 * BQ);
  const int units = p.n_qpairs * num_heads;
  const int sms = sm_count();
  if (sms <= 0) return SVI_ERR_DRIVER;
  SVI_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "%s: workspace must be 16-byte aligned", who);
  plan_split(units, (Lk + BKV - 1) / BKV, sms, (workspace && !accumulate) ? workspace_bytes : 0, &p.n_full, &p.split);
  const int n_sliced = units - p.n_full;
  const long long slices = (long long)n_sliced * p.split;
  p.ws_o = static_cast<float*>(workspace);
  p.ws_ml = reinterpret_cast<float2*>(static_cast<char*>(workspace) + (size_t)slices * 2 * BQ * HD * 4);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  attn_fwd_kernel<<<(unsigned)(p.n_full + slices), NUM_THREADS, SMEM_BYTES, st>>>(tq, tk, tv, p);
  SVI_CUDA_LAUNCH_CHECK(who);
  if (p.split > 1) {
    attn_merge_kernel<<<n_sliced * (2 * BQ / 8), 256, 0, st>>>(p.ws_o, p.ws_ml, p.split, p.n_full, p.n_qpairs, Lq, p.O, ldo);
    SVI_CUDA_LAUNCH_CHECK(who);
  }
  return SVI_OK;
}

}  // namespace attn
}  // namespace svi

extern "C" int svi_attn_fwd(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V,
                            int64_t ldv, void* O, int64_t ldo, int32_t Lq, int32_t Lk,
                            int32_t num_heads, float scale, int32_t accumulate, void* workspace,
                            size_t workspace_bytes, void* stream) {
  return svi::attn::launch_attn(Q, ldq, K, ldk, V, ldv, O, ldo, Lq, Lk, num_heads, scale, accumulate, nullptr, 0, 1, 0,
                                workspace, workspace_bytes, stream, "svi_attn_fwd");
}

extern "C" int svi_attn_fwd_qscale(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                   void* O, int64_t ldo, int32_t Lq, int32_t Lk, int32_t num_heads, float scale,
                                   int32_t accumulate, const float* q_sumsq, int32_t q_ss_ld, int32_t q_ss_parts, int32_t q_dim,
                                   float q_eps, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace svi;
  SVI_REQUIRE(q_sumsq && q_ss_ld >= 1 && q_dim >= 1 && q_ss_parts >= 0 && q_ss_parts <= q_ss_ld,
              "svi_attn_fwd_qscale: need q_sumsq, q_ss_ld >= 1, q_dim >= 1, 0 <= q_ss_parts <= q_ss_ld");
  svi::attn::QScale qs;
  qs.sumsq = q_sumsq;
  qs.ld = q_ss_ld;
  qs.parts = q_ss_parts > 0 ? q_ss_parts : 1;
  qs.dim = q_dim;
  qs.eps = q_eps;
  return svi::attn::launch_attn(Q, ldq, K, ldk, V, ldv, O, ldo, Lq, Lk, num_heads, scale, accumulate, nullptr, 0, 1, 0,
                                workspace, workspace_bytes, stream, "svi_attn_fwd_qscale", qs);
}

extern "C" void svi_attn_plan(int32_t units, int32_t kv_tiles, int32_t sms, size_t workspace_bytes, int32_t* n_full,
                              int32_t* split) {
  int a = units, b = 1;
  if (sms > 0) svi::attn::plan_split(units, kv_tiles, sms, workspace_bytes, &a, &b);
  if (n_full) *n_full = a;
  if (split) *split = b;
}

extern "C" size_t svi_attn_workspace_bytes(int32_t Lq, int32_t Lk, int32_t num_heads) {
  // enough for the largest plan plan_split can choose: up to two rounds of units, each cut into MAX_SPLIT slices
  using namespace svi::attn;
  (void)Lk;
  const long long units = (long long)((Lq + 2 * BQ - 1) / (2 * BQ)) * num_heads;
  const int sms = svi::sm_count();
  const long long tail = units < 2LL * sms ? units : 2LL * sms;
  return (size_t)tail * MAX_SPLIT * SLICE_BYTES;
}

extern "C" int svi_attn_fwd_sp(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V,
                               int64_t ldv, void* O, int64_t ldo, int32_t Lq, int32_t Lk,
                               int32_t num_heads, float scale, const void* kv_flags, uint32_t kv_epoch,
                               int32_t kv_chunk_rows, int32_t kv_self_chunk, void* workspace, size_t workspace_bytes,
                               void* stream) {
  using namespace svi;
  SVI_REQUIRE(kv_flags, "svi_attn_fwd_sp: kv_flags is null");
  SVI_REQUIRE(kv_chunk_rows >= 128 && kv_self_chunk >= 0 && (int64_t)kv_self_chunk * kv_chunk_rows < Lk,
              "svi_attn_fwd_sp: need kv_chunk_rows >= 128 and the local chunk inside [0, Lk)");
  return svi::attn::launch_attn(Q, ldq, K, ldk, V, ldv, O, ldo, Lq, Lk, num_heads, scale, 0,
                                static_cast<const uint32_t*>(kv_flags), kv_epoch, kv_chunk_rows, kv_self_chunk, workspace,
                                workspace_bytes, stream, "svi_attn_fwd_sp");
}
