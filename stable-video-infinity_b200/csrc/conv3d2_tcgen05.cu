// Causal 3-D / 2-D convolution of the Wan VAE on CTA PAIRS with reuse of the input window across the horizontal taps
// (tcgen05 cta_group::2, sm_100a).  Second kernel behind svi_conv3d_causal (conv3d_tcgen05.cu holds the C-ABI entry and the
// single-CTA kernel it started with).
//
// Why: the single-CTA kernel loads one 128-pixel x 64-channel input box AND one C_out x 64 weight box per (k_t, k_h, k_w,
// channel chunk).  For the layers that carry the VAE's FLOPs (C_in = C_out = 96 at full resolution, 192 at half) that is
// 40-120 KB of L2 -> shared-memory traffic per 384-1150 tensor-pipe cycles, 100-145 B/clk per SM against the ~60 B/clk the
// L2 delivers: the tensor pipe measured 36 % busy (profiles/r02_conv_ncu_summary.txt).  Here
//   * a CTA's tile is 128 consecutive pixels of ONE image row; the input box is loaded once per (k_t, k_h, chunk) with
//     k_w - 1 extra pixels (136 rows of 128 B, 128B-swizzled by TMA) and the k_w taps are k_w MMAs whose A descriptors start
//     0, 1, 2 rows into that box — input traffic / 3;
//   * two CTAs (image rows h, h+1) form a cluster: one tcgen05.mma.cta_group::2 (M = 256) consumes both CTAs' input rows
//     and a weight tile of which each CTA loaded only half — weight traffic / 2;
//   * the last channel chunk issues only the K-steps that hold real channels (C_in = 96: 2 of 4).
// Stage = input box + k_w weight half-tiles (35 KB at C_out = 96: 6 stages, 53 KB at 192: 4 stages).
// Roles per CTA (384 threads, role functions on 32-bit shared addresses so that setmaxnreg can move registers): warps 0-7
// epilogue (two per TMEM lane quadrant, half of the output channels each; the pixel's sum of squares for the fused RMS norm is
// exchanged through shared memory; the pixel's fp32 residual waits in registers from before the accumulator is ready), warp 8
// TMA producer, warp 9 MMA issuer (even CTA) + TMEM owner, warps 10-11 idle.  Epilogue semantics are those of conv3d_tcgen05.cu.
#include "common.cuh"
#include "../../include/svi_b200.h"
#include "conv3d_common.cuh"

namespace svi {
namespace conv2 {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int AW = 136;                       // input pixels per box: 128 + (k_w - 1), rounded up to whole 8-row swizzle groups
constexpr int A_BYTES = AW * BK * 2;          // 17 KB
constexpr int MAX_BN = 256;
constexpr int MAX_STAGES = 8;
constexpr int RING_BYTES = 216 * 1024;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 384;              // 3 warpgroups: warps 0-7 epilogue, 8 TMA, 9 MMA + TMEM owner, 10-11 idle (setmaxnreg)
constexpr int TMEM_COLS = 512;
constexpr int SSQ_BYTES = 2 * 2 * BM * 4;     // [accumulator][column half][pixel]
constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256 + SSQ_BYTES;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads whose completion bytes are credited to the EVEN CTA's mbarrier
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_even, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_dst),
      "l"(m), "r"(bar_even), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_even, int c_inner,
                                                int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(m), "r"(bar_even), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// All MMAs of one operand stage — 3 horizontal taps x NK K = 16 steps of the M = 256 pair MMA — in ONE asm block (whole warp
// calls with uniform operands; one elected lane issues).  The MMA warp is issue-bound at N = 96: an MMA occupies the tensor
// pipe for 48 cycles, and with a C++ loop of single MMAs the warp spent ~24 instructions (~130 cycles) per MMA on elects,
// predicates, descriptor arithmetic and register -> uniform-register moves (ncu: the warp never waited for operands while the
// tensor pipe was 31-36 % busy, profiles/r02_c7 / r02_c8).  Here an MMA costs its descriptor adds, four R2UR and the UTCHMMA.
// Tap c reads the A box c rows (c * 128 B) further in and weight tile c (b_step 16-byte units apart); K-step k is 32 B on.
#define SVI_MMA2_OPEN                      \
  "{\n"                                    \
  ".reg .pred p, q, t;\n"                  \
  ".reg .b64 da, db;\n"                    \
  ".reg .b32 a1, b1, bt;\n"                \
  "setp.ne.b32 p, %6, 0;\n"                \
  "setp.eq.b32 t, 0, 0;\n"                 \
  "elect.sync _|q, 0xffffffff;\n"          \
  "mov.b32 bt, %3;\n"
#define SVI_MMA2_ONE(aoff, koff, pred)     \
  "add.u32 a1, %1, " #aoff ";\n"           \
  "add.u32 b1, bt, " #koff ";\n"           \
  "mov.b64 da, {a1, %2};\n"                \
  "mov.b64 db, {b1, %4};\n"                \
  "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, " #pred ";\n"
#define SVI_MMA2_NEXT_TAP "add.u32 bt, bt, %7;\n"
#define SVI_MMA2_ARGS \
  ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first), "r"(b_step) : "memory"
// A offsets: tap c adds 8 (128 B >> 4), K-step k adds 2 (32 B >> 4)
#define SVI_MMA2_TAP4(c, first) SVI_MMA2_ONE(c * 8 + 0, 0, first) SVI_MMA2_ONE(c * 8 + 2, 2, t) SVI_MMA2_ONE(c * 8 + 4, 4, t) SVI_MMA2_ONE(c * 8 + 6, 6, t)
#define SVI_MMA2_TAP2(c, first) SVI_MMA2_ONE(c * 8 + 0, 0, first) SVI_MMA2_ONE(c * 8 + 2, 2, t)
#define SVI_MMA2_TAP1(c, first) SVI_MMA2_ONE(c * 8 + 0, 0, first)
#define SVI_MMA2_TAP3(c, first) SVI_MMA2_ONE(c * 8 + 0, 0, first) SVI_MMA2_ONE(c * 8 + 2, 2, t) SVI_MMA2_ONE(c * 8 + 4, 4, t)
__device__ __forceinline__ void mma2_stage(int nk, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t b_step, uint32_t idesc, uint32_t acc_first) {
  if (nk == 4)
    asm volatile(SVI_MMA2_OPEN SVI_MMA2_TAP4(0, p) SVI_MMA2_NEXT_TAP SVI_MMA2_TAP4(1, t) SVI_MMA2_NEXT_TAP SVI_MMA2_TAP4(2, t) "}\n" SVI_MMA2_ARGS);
  else if (nk == 2)
    asm volatile(SVI_MMA2_OPEN SVI_MMA2_TAP2(0, p) SVI_MMA2_NEXT_TAP SVI_MMA2_TAP2(1, t) SVI_MMA2_NEXT_TAP SVI_MMA2_TAP2(2, t) "}\n" SVI_MMA2_ARGS);
  else if (nk == 1)
    asm volatile(SVI_MMA2_OPEN SVI_MMA2_TAP1(0, p) SVI_MMA2_NEXT_TAP SVI_MMA2_TAP1(1, t) SVI_MMA2_NEXT_TAP SVI_MMA2_TAP1(2, t) "}\n" SVI_MMA2_ARGS);
  else
    asm volatile(SVI_MMA2_OPEN SVI_MMA2_TAP3(0, p) SVI_MMA2_NEXT_TAP SVI_MMA2_TAP3(1, t) SVI_MMA2_NEXT_TAP SVI_MMA2_TAP3(2, t) "}\n" SVI_MMA2_ARGS);
}
#undef SVI_MMA2_OPEN
#undef SVI_MMA2_ONE
#undef SVI_MMA2_NEXT_TAP
#undef SVI_MMA2_ARGS
#undef SVI_MMA2_TAP4
#undef SVI_MMA2_TAP3
#undef SVI_MMA2_TAP2
#undef SVI_MMA2_TAP1
__device__ __forceinline__ void commit2_multicast(uint32_t bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      ".reg .b16 m;\n"
      "mov.b16 m, 3;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n"
      "}\n" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t target_rank) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar),
      "r"(target_rank)
      : "memory");
}
#ifndef SVI_CONV2_UNIFORM
#define SVI_CONV2_UNIFORM 1     // MMA operands as shuffle-from-lane-0 values (uniform datapath); 0: plain registers (A/B: +1 % with 1)
#endif
#ifndef SVI_CONV2_EPI
#define SVI_CONV2_EPI 0         // 0: the residual waits in registers, both epilogue passes read the accumulator; 1: the first pass leaves
                                // the result in registers and frees the accumulator, the second works from registers.  Measured
                                // (profiles/r02_c11_conv_variants.log, TF/s 0 / 1): C = 96 'mid' 881 / 793, 'end' 603 / 546; C = 192 'mid'
                                // 1284 / 1330, 'end' 898 / 922; whole VAE decode 382.4 / 384.0 ms -> 0
#endif
__device__ __forceinline__ uint32_t uniform(uint32_t v) {
#if SVI_CONV2_UNIFORM
  return __shfl_sync(0xffffffffu, v, 0);
#else
  return v;
#endif
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct Params {
  svi::conv::Params c;          // geometry + epilogue of the single-CTA kernel (BW / BH / stages unused here)
  int stage_bytes, stages;
  int b_bytes;                  // one weight half-tile: (BN / 2) rows x 128 B
  int nk_last;                  // K = 16 steps of the last channel chunk that hold real channels
};

// Register budget: 384 threads start with 168 registers; once the roles are fixed the TMA / MMA / idle warpgroup drops to 88
// and the two epilogue warpgroups take 208 (2 * 208 + 88 = 3 * 168: setmaxnreg only redistributes what the CTA owns).  The
// epilogue needs them to hold a pixel's whole fp32 residual (up to 128 channels per warp) from before the accumulator is
// ready until the second pass.
constexpr int EPI_REGS = 208, OTHER_REGS = 88;
static_assert(2 * EPI_REGS + OTHER_REGS <= 3 * 168 && EPI_REGS % 8 == 0 && OTHER_REGS % 8 == 0, "register split");

enum : uint32_t { FULL = 0, EMPTY = MAX_STAGES, TMEM_FULL = 2 * MAX_STAGES, TMEM_EMPTY = 2 * MAX_STAGES + 2, NUM_BARS = 2 * MAX_STAGES + 4 };
__device__ __forceinline__ uint32_t bar_addr(uint32_t sbase, uint32_t n) { return sbase + RING_BYTES + 8u * n; }

struct TileWalk {            // tile -> (n block fastest, then 128-pixel segment, then row pair, then frame)
  int tiles_w, tiles_per_frame, num_n, num_tiles;
  __device__ __forceinline__ TileWalk(const svi::conv::Params& p) {
    tiles_w = (p.W + BM - 1) / BM;
    tiles_per_frame = ((p.H + 1) / 2) * tiles_w;
    num_n = (p.C_out + p.BN - 1) / p.BN;
    num_tiles = p.T * tiles_per_frame * num_n;
  }
  __device__ __forceinline__ void decode(int tile, int& t, int& h0, int& w0, int& n_blk) const {
    n_blk = tile % num_n;
    const int r = tile / num_n;
    const int sp = r % tiles_per_frame;
    t = r / tiles_per_frame;
    h0 = (sp / tiles_w) * 2;
    w0 = (sp % tiles_w) * BM;
  }
};

__device__ __forceinline__ void producer_role(const CUtensorMap* tmap_x, const CUtensorMap* tmap_w, const Params& pp, uint32_t sbase,
                                              uint32_t rank, int pair, int num_pairs) {
  const svi::conv::Params& p = pp.c;
  const TileWalk tw(p);
  const int BNH = p.BN / 2;
  int stage = 0;
  uint32_t phase = 0;
  const uint32_t stage_tx = A_BYTES + p.kw * pp.b_bytes;
  for (int tile = pair; tile < tw.num_tiles; tile += num_pairs) {
    int t, h0, w0, n_blk;
    tw.decode(tile, t, h0, w0, n_blk);
    const int h = h0 + (int)rank;
    const int wrow = n_blk * p.BN + (int)rank * BNH;
    for (int a = 0; a < p.kt; ++a) {
      const int slot = p.slot[t][a];
      for (int b = 0; b < p.kh; ++b) {
        for (int cc = 0; cc < p.cin_chunks; ++cc) {
          mbar_wait_a(bar_addr(sbase, EMPTY + stage), phase ^ 1);
          if (rank == 0) mbar_expect_tx_a(bar_addr(sbase, FULL + stage), 2 * stage_tx);
          const uint32_t sa = sbase + stage * pp.stage_bytes;
          const uint32_t full_even = bar_addr(sbase, FULL + stage) & PEER_MASK;
          tma_load_4d_2sm(sa, tmap_x, full_even, cc * BK, w0 - p.pad_w, h + b - p.pad_h, slot);
          const int kcol0 = ((a * p.kh + b) * p.kw * p.cin_chunks + cc) * BK;
          for (int c = 0; c < p.kw; ++c)
            tma_load_2d_2sm(sa + A_BYTES + c * pp.b_bytes, tmap_w, full_even, kcol0 + c * p.cin_chunks * BK, wrow);
          if (++stage == pp.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  }
}

__device__ __forceinline__ void mma_role(const Params& pp, uint32_t sbase, uint32_t tmem_base, int pair, int num_pairs) {
  const svi::conv::Params& p = pp.c;
  const TileWalk tw(p);
  // operands of the MMA asm block as provably warp-uniform values (a shuffle from lane 0): the descriptor arithmetic then runs
  // in the uniform datapath and UTCHMMA takes its uniform registers directly instead of through R2UR moves
  const uint32_t idesc = uniform(make_idesc_bf16(2 * BM, p.BN, 0, 0));
  const uint32_t b_step = uniform((uint32_t)pp.b_bytes >> 4);
  constexpr uint32_t hi_kmaj = smem_desc_hi(1024, 2);  // SBO 1024 B, 128B swizzle, K-major
  int stage = 0;
  uint32_t phase = 0;
  int acc = 0;
  uint32_t acc_phase = 0;
  const int groups = p.kt * p.kh;
  for (int tile = pair; tile < tw.num_tiles; tile += num_pairs) {
    mbar_wait_a(bar_addr(sbase, TMEM_EMPTY + acc), acc_phase ^ 1);
    tc_fence_after();
    const uint32_t d_tmem = uniform(tmem_base + acc * MAX_BN);
    uint32_t accumulate = 0;
    for (int g = 0; g < groups; ++g) {
      for (int cc = 0; cc < p.cin_chunks; ++cc) {
        mbar_wait_a(bar_addr(sbase, FULL + stage), phase);
        tc_fence_after();
        const uint32_t sa = sbase + stage * pp.stage_bytes;
        const int nk = cc == p.cin_chunks - 1 ? pp.nk_last : 4;
        // tap c of this row = the same box read c pixels (rows of 128 B) further in.  The 128B swizzle is a function of the
        // absolute shared-memory address (TMA wrote the box into a 1024-byte aligned stage), so a start address that is not a
        // multiple of 8 rows needs nothing else: the descriptor's base-offset field stays 0 (measured on B200: with the row
        // phase in that field the results are wrong, profiles/r02_c6_pair_probe.log).  k_w == 3 (launch side).
        mma2_stage(nk, d_tmem, uniform(smem_desc_lo(sa, 16)), hi_kmaj, uniform(smem_desc_lo(sa + A_BYTES, 16)), hi_kmaj, b_step,
                   idesc, accumulate);
        accumulate = 1;
        commit2_multicast(bar_addr(sbase, EMPTY + stage));
        if (++stage == pp.stages) { stage = 0; phase ^= 1; }
      }
    }
    commit2_multicast(bar_addr(sbase, TMEM_FULL + acc));
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
}

// Epilogue of one warp: TMEM lane quadrant `quad`, output channels [half * BN/2, (half + 1) * BN/2) of the tile.  A pixel's
// fp32 residual (strided 16-byte reads, one line per lane: ~1-2 k cycles each) is fetched into registers BEFORE the accumulator
// is waited for, i.e. under the K loop of the tile, and serves both passes (with loads inside the passes the 'end' convolutions of
// the ResidualBlocks ran 30-45 % slower than the 'mid' ones: profiles/r02_c8_perf_conv.log, r02_c7_conv2_192_end ncu).
constexpr int MAX_Q = MAX_BN / 2 / 4;     // float4 per pixel and warp: 32
__device__ __forceinline__ void epilogue_role(const Params& pp, uint32_t sbase, uint32_t tmem_base, float* ssq_x, uint32_t rank,
                                              int pair, int num_pairs, int warp, int lane) {
  const svi::conv::Params& p = pp.c;
  const TileWalk tw(p);
  const int quad = warp & 3;
  const int half = warp >> 2;
  const int row_in_tile = quad * 32 + lane;
  const int cols_half = p.BN / 2;           // multiple of 16 (launch side)
  const int n4 = cols_half / 4;
  int acc = 0;
  uint32_t acc_phase = 0;
  for (int tile = pair; tile < tw.num_tiles; tile += num_pairs) {
    int t, h0, w0, n_blk;
    tw.decode(tile, t, h0, w0, n_blk);
    const int h = h0 + (int)rank, w = w0 + row_in_tile;
    const bool ok = (h < p.H) && (w < p.W);
    const long long pix = (long long)h * p.W + w;
    const int col0 = n_blk * p.BN + half * cols_half;
    float4 q[MAX_Q];
    {
      const float* rrow = p.residual ? p.residual + (long long)t * p.res_frame_stride + pix * p.res_ld + col0 : nullptr;
#pragma unroll
      for (int j = 0; j < MAX_Q; ++j)
        q[j] = (rrow && ok && j < n4 && col0 + j * 4 < p.C_out) ? *reinterpret_cast<const float4*>(rrow + j * 4)
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    mbar_wait_a(bar_addr(sbase, TMEM_FULL + acc), acc_phase);
    tc_fence_after();
    const uint32_t t_base = tmem_base + acc * MAX_BN + half * cols_half + (static_cast<uint32_t>(quad * 32) << 16);
#if SVI_CONV2_EPI == 1
    // pass 1: v = acc + bias + residual -> fp32 out, sum of squares; v replaces the residual in the registers
    float ssq = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_Q / 4; ++c) {           // 16 columns per step
      const int n0 = col0 + c * 16;
      if (c * 16 < cols_half && n0 < p.C_out) {      // warp-uniform
        uint32_t r[16];
        tmem_ld16(t_base + c * 16, r);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int n = n0 + j4 * 4;
            if (n < p.C_out) {
              float4 v = q[c * 4 + j4];
              v.x += __uint_as_float(r[j4 * 4 + 0]); v.y += __uint_as_float(r[j4 * 4 + 1]);
              v.z += __uint_as_float(r[j4 * 4 + 2]); v.w += __uint_as_float(r[j4 * 4 + 3]);
              if (p.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
              }
              ssq += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
              q[c * 4 + j4] = v;
              if (p.write_f32) {
                float* dst;
                if (p.n_split > 0 && n >= p.n_split)
                  dst = p.out + p.split_offset + (long long)t * p.out_frame_stride + pix * p.out_ld + (n - p.n_split);
                else
                  dst = p.out + (long long)t * p.out_frame_stride + pix * p.out_ld + n;
                *reinterpret_cast<float4*>(dst) = v;
              }
            }
          }
        }
      }
    }
    // the accumulator is free from here on: the second pass works on the registers
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(bar_addr(sbase, TMEM_EMPTY + acc), 0);
    if (p.next_ring) {
      if (p.next_gamma) {       // the other half of the pixel's channel vector belongs to the partner warp of this quadrant
        float* sx = ssq_x + acc * 2 * BM;
        sx[half * BM + row_in_tile] = ssq;
        named_bar_sync(1 + quad, 64);
        ssq = sx[row_in_tile] + sx[BM + row_in_tile];
      }
      // pass 2: RMS norm + SiLU of v -> bf16 into the next conv's ring
      const float mul = p.next_gamma ? sqrtf((float)p.C_out) / fmaxf(sqrtf(ssq), 1e-12f) : 1.f;
      __nv_bfloat16* nrow = p.next_ring + (long long)p.next_slot[t] * p.next_frame_stride + pix * p.next_ld;
      if (ok) {
#pragma unroll
        for (int j = 0; j < MAX_Q; ++j) {
          const int n = col0 + j * 4;
          if (j < n4 && n < p.C_out) {
            float4 v = q[j];
            if (p.next_gamma) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(p.next_gamma + n));
              v.x *= mul * g.x; v.y *= mul * g.y; v.z *= mul * g.z; v.w *= mul * g.w;
            }
            if (p.next_silu) { v.x = silu(v.x); v.y = silu(v.y); v.z = silu(v.z); v.w = silu(v.w); }
            uint2 pk;
            pk.x = pack_bf16x2(v.x, v.y);
            pk.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(nrow + n) = pk;
          }
        }
      }
    }
#else   // A/B: residual in registers, both passes read the accumulator, released at the end of the tile
    // pass 1: v = acc + bias + residual -> fp32 out, sum of squares
    float ssq = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_Q / 4; ++c) {           // 16 columns per step
      const int n0 = col0 + c * 16;
      if (c * 16 < cols_half && n0 < p.C_out) {      // warp-uniform
        uint32_t r[16];
        tmem_ld16(t_base + c * 16, r);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int n = n0 + j4 * 4;
            if (n < p.C_out) {
              float v[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r[j4 * 4 + j]);
              if (p.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
              }
              const float4 qq = q[c * 4 + j4];
              v[0] += qq.x; v[1] += qq.y; v[2] += qq.z; v[3] += qq.w;
              ssq += v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
              if (p.write_f32) {
                float* dst;
                if (p.n_split > 0 && n >= p.n_split)
                  dst = p.out + p.split_offset + (long long)t * p.out_frame_stride + pix * p.out_ld + (n - p.n_split);
                else
                  dst = p.out + (long long)t * p.out_frame_stride + pix * p.out_ld + n;
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
              }
            }
          }
        }
      }
    }
    if (p.next_ring) {
      if (p.next_gamma) {       // the other half of the pixel's channel vector belongs to the partner warp of this quadrant
        float* sx = ssq_x + acc * 2 * BM;
        sx[half * BM + row_in_tile] = ssq;
        named_bar_sync(1 + quad, 64);
        ssq = sx[row_in_tile] + sx[BM + row_in_tile];
      }
      // pass 2 (accumulator still in TMEM, residual still in registers): RMS norm + SiLU -> bf16 into the next conv's ring
      const float mul = p.next_gamma ? sqrtf((float)p.C_out) / fmaxf(sqrtf(ssq), 1e-12f) : 1.f;
      __nv_bfloat16* nrow = p.next_ring + (long long)p.next_slot[t] * p.next_frame_stride + pix * p.next_ld;
#pragma unroll
      for (int c = 0; c < MAX_Q / 4; ++c) {
        const int n0 = col0 + c * 16;
        if (c * 16 < cols_half && n0 < p.C_out) {
          uint32_t r[16];
          tmem_ld16(t_base + c * 16, r);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const int n = n0 + j4 * 4;
              if (n < p.C_out) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r[j4 * 4 + j]);
                if (p.bias) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
                  v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
                }
                const float4 qq = q[c * 4 + j4];
                v[0] += qq.x; v[1] += qq.y; v[2] += qq.z; v[3] += qq.w;
                if (p.next_gamma) {
                  const float4 g = __ldg(reinterpret_cast<const float4*>(p.next_gamma + n));
                  v[0] *= mul * g.x; v[1] *= mul * g.y; v[2] *= mul * g.z; v[3] *= mul * g.w;
                }
                if (p.next_silu) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) v[j] = silu(v[j]);
                }
                uint2 pk;
                pk.x = pack_bf16x2(v[0], v[1]);
                pk.y = pack_bf16x2(v[2], v[3]);
                *reinterpret_cast<uint2*>(nrow + n) = pk;
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(bar_addr(sbase, TMEM_EMPTY + acc), 0);
#endif
    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
             const __grid_constant__ Params pp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  const uint32_t tmem_slot = sbase + RING_BYTES + 8 * NUM_BARS;
  float* ssq_x = reinterpret_cast<float*>(smem_raw + (sbase - sraw) + RING_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 9) {
    if (lane == 0) {
      for (uint32_t i = 0; i < MAX_STAGES; ++i) {
        mbar_init_a(bar_addr(sbase, FULL + i), 1);
        mbar_init_a(bar_addr(sbase, EMPTY + i), 1);
      }
      for (uint32_t i = 0; i < 2; ++i) {
        mbar_init_a(bar_addr(sbase, TMEM_FULL + i), 1);
        mbar_init_a(bar_addr(sbase, TMEM_EMPTY + i), 2 * EPI_WARPS);
      }
      fence_mbar_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp >= 8) {          // warps 10-11 only complete the third warpgroup (setmaxnreg is warpgroup-wide)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(OTHER_REGS));
    if (warp == 8) {
      if (lane == 0) producer_role(&tmap_x, &tmap_w, pp, sbase, rank, pair, num_pairs);
    } else if (warp == 9) {
      if (rank == 0) mma_role(pp, sbase, tmem_base, pair, num_pairs);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    epilogue_role(pp, sbase, tmem_base, ssq_x, rank, pair, num_pairs, warp, lane);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// the launch side: validated descriptor + parameters come from svi_conv3d_causal (conv3d_tcgen05.cu)

// N tile of the pair kernel for this convolution (a multiple of 32: each CTA loads half of it in 8-row swizzle groups and each
// of the two epilogue warps of a lane quadrant drains a multiple of 16 columns), or 0 when the kernel does not apply.
// Narrow outputs (the decoder head, 3 channels padded to 4) run with N = 32: weight rows past w_rows are zero-filled by TMA.
int pair_bn(const svi_conv_desc* d, int BN) {
  if (d->kw != 3 || d->pad_w != 1) return 0;
  if (d->C_out <= MAX_BN) return (d->C_out + 31) / 32 * 32;
  return BN % 32 == 0 && BN <= MAX_BN ? BN : 0;
}

int launch(const svi_conv_desc* d, const svi::conv::Params& base, int BN, cudaStream_t stream) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(svi::conv::encode_tiled_fn());
  if (!enc) return SVI_ERR_DRIVER;
  CUtensorMap tx, tw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->C_in, (cuuint64_t)d->in_W, (cuuint64_t)d->in_H, (cuuint64_t)d->ring_slots};
    cuuint64_t strides[3] = {(cuuint64_t)d->C_in * 2, (cuuint64_t)d->in_W * d->C_in * 2,
                             (cuuint64_t)d->in_H * d->in_W * d->C_in * 2};
    cuuint32_t box[4] = {BK, (cuuint32_t)AW, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x_ring), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("svi_conv3d_causal(pair): input tensor map failed (CUresult %d) C=%d W=%d H=%d slots=%d", (int)r, d->C_in,
                     d->in_W, d->in_H, d->ring_slots);
      return SVI_ERR_DRIVER;
    }
  }
  const int cin_chunks = (d->C_in + BK - 1) / BK;
  const int ktot = d->kt * d->kh * d->kw * cin_chunks * BK;
  int rc = make_tmap_2d(&tw, d->w_packed, 2, (uint64_t)ktot, (uint64_t)d->w_rows, (uint64_t)d->w_ld * 2, BK, BN / 2);
  if (rc) return rc;

  Params pp;
  pp.c = base;
  pp.c.BN = BN;
  pp.b_bytes = (BN / 2) * BK * 2;
  pp.stage_bytes = A_BYTES + d->kw * pp.b_bytes;
  pp.stages = RING_BYTES / pp.stage_bytes;
  if (pp.stages > MAX_STAGES) pp.stages = MAX_STAGES;
  SVI_REQUIRE(pp.stages >= 2, "svi_conv3d_causal(pair): stage of %d bytes does not fit twice", pp.stage_bytes);
  const int last = d->C_in - (cin_chunks - 1) * BK;
  pp.nk_last = (last + 15) / 16;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(conv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce != cudaSuccess) {
      set_last_error("svi_conv3d_causal(pair): cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce));
      return SVI_ERR_LAUNCH;
    }
    attr_set = true;
  }
  const int tiles = d->T * ((d->H + 1) / 2) * ((d->W + BM - 1) / BM) * ((d->C_out + BN - 1) / BN);
  const int sms = sm_count();
  if (sms <= 0) return SVI_ERR_DRIVER;
  int pairs = sms / 2;
  if (tiles < pairs) pairs = tiles;
  conv2_kernel<<<2 * pairs, NUM_THREADS, SMEM_BYTES, stream>>>(tx, tw, pp);
  SVI_CUDA_LAUNCH_CHECK("svi_conv3d_causal(pair)");
  return SVI_OK;
}

}  // namespace conv2
}  // namespace svi
