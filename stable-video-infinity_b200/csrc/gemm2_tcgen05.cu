// bf16 GEMM on CTA PAIRS (tcgen05 cta_group::2):  out = epilogue(A[M,K] @ W[N,K]^T), tile 256 x 256 x 64 per pair.
//
// Why: the single-CTA kernel (gemm_tcgen05.cu, 128x256 tiles) moves 48 KB of TMA writes + 48 KB of MMA operand reads
// through one SM's shared memory per 512 tensor cycles (192 B/clk against a ~128 B/clk port) and measures 85 % of
// cuBLAS.  In a CTA pair every CTA holds its own 128 rows of A and only HALF of the 256-row W tile; one
// tcgen05.mma.cta_group::2 (M=256) issued by the even CTA drives both SMs' tensor cores, each reading its local A
// and both halves of W.  Per SM: 32 KB written + 4 KB(A) + 4 KB(W half) read per K-step -> 128 B/clk.
//
//   cluster = 2 CTAs (same TPC).  Per CTA: warp 0 TMA producer (its A rows + its half of W; completion bytes of
//   BOTH CTAs land on the even CTA's `full` barrier), warp 1 MMA issuer (even CTA only) + TMEM owner
//   (cta_group::2 allocation), warps 2-9 epilogue of the CTA's own 128 accumulator rows (two warps per TMEM lane
//   quadrant, 128 columns each: with activation / residual epilogues four warps took longer than the K loop of a
//   K = 1536 tile and the tensor pipe waited for a free accumulator).
//   `empty` / `tmem_full` are signalled in both CTAs by multicast commits; the odd CTA's epilogue releases the
//   accumulator with a remote arrive on the even CTA's `tmem_empty`.
// Epilogue and C-ABI contract are those of gemm_tcgen05.cu (shared entry point svi_gemm_bf16).
#include "common.cuh"
#include "../../include/svi_b200.h"
#include "gemm_epilogue.cuh"

namespace svi {
namespace gemm2 {

constexpr int BM = 128;         // rows per CTA (256 per pair)
constexpr int BN = 256;
constexpr int BN_HALF = 128;    // W rows loaded by each CTA
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KB
constexpr int B_STAGE_BYTES = BN_HALF * BK * 2;   // 16 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int EPI_WARPS = 8;     // two warps per TMEM lane quadrant, each draining one 128-column half of the tile
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int TMEM_COLS = 512;
constexpr int MAX_STAGES = 6;   // operand ring: 6 stages (7 fit but measured 1-2 % slower); 5 when the residual-class epilogue needs its
                                // transposition tiles in the same 227 KB (a runtime parameter: ffn.0 lost 16 % with a 5-stage ring)
constexpr int GROUP_M = 4;      // measured: 1..4 equal on M >> N shapes, 4..8 best on square ones
constexpr int VEC_BYTES = 2 * 4 * BN * 4;   // bias, gate, LN-fold u and next-operand scale of the tile's 256 columns, double buffered by accumulator
constexpr int XPOSE_BYTES = EPI_WARPS * 32 * 32 * 4;   // one 32 x 32 fp32 tile per epilogue warp (residual-class epilogue)
constexpr int smem_bytes(int stages, bool xpose) { return stages * STAGE_BYTES + 1024 + 256 + VEC_BYTES + (xpose ? XPOSE_BYTES : 0); }
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> even CTA

struct Epi {
  void* out;
  long long ldo;
  int out_is_f32;
  int act;
  const float* bias;
  const float* gate;
  const float* residual;
  long long ldr;
  float* sumsq;
  int sumsq_groups;
  int sumsq_group_cols;
  int sumsq_parts;
  // LayerNorm folded into this GEMM (consumer side): A holds bf16(x * g) of the UN-normalised rows x; with the row statistics
  // (sum x, sum x^2) the epilogue reconstructs  LN(x)*g + t  times W^T  =  r_m (acc - mu_m u_n) + c_n,  u = W g, c = W t + b
  // (c arrives as `bias`)
  const float* ln_stats;   // [M, 2] or null
  const float* ln_u;       // [N]
  float ln_inv_d, ln_eps;
  // producer side: besides out, emit the NEXT GEMM's folded operand and row statistics from the final value v
  __nv_bfloat16* a_next;   // [M, ld_an] = bf16(v * g_next[n]) or null
  long long ld_an;
  const float* g_next;     // [N]
  float* row_stats;        // [M, 2] += (sum v, sum v^2) over this launch's columns (atomicAdd)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are credited to the EVEN CTA's mbarrier (address with the rank bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_even, int c_inner,
                                                int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_dst),
      "l"(m), "r"(bar_even), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// four K-steps of the M=256 pair MMA (issued by one elected lane of the even CTA's MMA warp)
__device__ __forceinline__ void mma2_ss_k4(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n"
      ".reg .pred p, q, t;\n"
      ".reg .b64 da, db;\n"
      ".reg .b32 a1, b1;\n"
      "setp.ne.b32 p, %6, 0;\n"
      "setp.eq.b32 t, 0, 0;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
      "add.u32 a1, %1, 2;\n"
      "add.u32 b1, %3, 2;\n"
      "mov.b64 da, {a1, %2};\n"
      "mov.b64 db, {b1, %4};\n"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, t;\n"
      "add.u32 a1, %1, 4;\n"
      "add.u32 b1, %3, 4;\n"
      "mov.b64 da, {a1, %2};\n"
      "mov.b64 db, {b1, %4};\n"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, t;\n"
      "add.u32 a1, %1, 6;\n"
      "add.u32 b1, %3, 6;\n"
      "mov.b64 da, {a1, %2};\n"
      "mov.b64 db, {b1, %4};\n"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, t;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first)
      : "memory");
}
// completion of all previously issued pair MMAs -> arrive on the barrier at this offset in BOTH CTAs
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
// arrive on the barrier at this shared::cta offset in CTA `target_rank` of the cluster.  Relaxed: the only thing the
// MMA issuer must observe is that this warp's tcgen05.ld have completed, which tcgen05.wait::ld + the tcgen05 fence
// before the arrive already order; a release here costs a MEMBAR over all the epilogue's global stores per tile.
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

// tile order: m-tiles are taken in groups of `group_m`; inside a group the m index runs fastest, so the pairs that
// are resident together cover a near-square patch of the output and share both A and W tiles in L2
__device__ __forceinline__ void tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk, int& n_blk) {
  const int per_group = group_m * num_n;
  const int g = tile / per_group;
  const int first_m = g * group_m;
  const int rows = min(group_m, num_m - first_m);
  const int idx = tile - g * per_group;
  m_blk = first_m + idx % rows;
  n_blk = idx / rows;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, int M, int N,
                  int K, int group_m, int STAGES, Epi ep) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  const uint32_t OFF_BAR = STAGES * STAGE_BYTES;
  enum : uint32_t { FULL = 0, EMPTY = MAX_STAGES, TMEM_FULL = 2 * MAX_STAGES, TMEM_EMPTY = 2 * MAX_STAGES + 2, NUM_BARS = 2 * MAX_STAGES + 4 };
  auto bar = [&](uint32_t n) { return sbase + OFF_BAR + 8u * n; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_raw + (sbase - sraw) + OFF_BAR + 8 * NUM_BARS);
  float* vec_smem = reinterpret_cast<float*>(smem_raw + (sbase - sraw) + OFF_BAR + 256);   // [2 acc][bias | gate | u | g_next][BN]
  uint8_t* xpose_smem = smem_raw + (sbase - sraw) + OFF_BAR + 256 + VEC_BYTES;            // [EPI_WARPS][32 rows][128 B], 16 B chunks XOR-swizzled

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // 0 = even CTA (MMA leader), 1 = odd CTA
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_m = (M + 2 * BM - 1) / (2 * BM);
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (uint32_t i = 0; i < MAX_STAGES; ++i) {
        mbar_init_a(bar(FULL + i), 1);      // one expect_tx arrive by the even CTA's producer (+ bytes of both CTAs)
        mbar_init_a(bar(EMPTY + i), 1);     // one multicast commit
      }
      for (uint32_t i = 0; i < 2; ++i) {
        mbar_init_a(bar(TMEM_FULL + i), 1);
        mbar_init_a(bar(TMEM_EMPTY + i), 2 * EPI_WARPS);  // epilogue warps of both CTAs (only the even CTA's copy is used)
      }
      fence_mbar_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation of both CTAs visible before anyone signals a peer
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------- TMA producer (both CTAs) -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int m_blk, n_blk;
        tile_coords(tile, num_m, num_n, group_m, m_blk, n_blk);
        const int row_a = m_blk * 2 * BM + (int)rank * BM;
        const int row_b = n_blk * BN + (int)rank * BN_HALF;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait_a(bar(EMPTY + stage), phase ^ 1);                                  // local: slot free in MY smem
          if (rank == 0) mbar_expect_tx_a(bar(FULL + stage), 2 * STAGE_BYTES);         // bytes of both CTAs
          const uint32_t sa = sbase + stage * STAGE_BYTES;
          const uint32_t full_even = bar(FULL + stage) & PEER_MASK;
          tma_load_2d_2sm(sa, &tmap_a, full_even, kb * BK, row_a);
          tma_load_2d_2sm(sa + A_STAGE_BYTES, &tmap_b, full_even, kb * BK, row_b);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (even CTA only) -----------------------------
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, 0, 0);   // M = 256 across the pair
      constexpr uint32_t hi_kmaj = smem_desc_hi(1024, 2);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait_a(bar(TMEM_EMPTY + acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait_a(bar(FULL + stage), phase);
          tc_fence_after();
          const uint32_t a_lo = smem_desc_lo(sbase + stage * STAGE_BYTES, 16);
          const uint32_t b_lo = a_lo + (A_STAGE_BYTES >> 4);
          mma2_ss_k4(d_tmem, a_lo, hi_kmaj, b_lo, hi_kmaj, idesc, kb != 0);
          commit2_multicast(bar(EMPTY + stage));                          // frees the slot in both CTAs
          if (kb == num_k - 1) commit2_multicast(bar(TMEM_FULL + acc));   // accumulators ready in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------- epilogue (both CTAs, own 128 rows) ---------------------
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may read (hardware rule: warp id % 4)
    const int col_half = (warp - 2) >> 2;       // which 128 accumulator columns this warp drains
    const int row_in_tile = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int m_blk, n_blk;
      tile_coords(tile, num_m, num_n, group_m, m_blk, n_blk);
      const int row = m_blk * 2 * BM + (int)rank * BM + row_in_tile;
      const bool row_ok = row < M;
      if (ep.residual && !(ep.out_is_f32 && ep.act == SVI_ACT_NONE && !ep.sumsq && !ep.ln_stats)) {
        // Residual rows -> L2 ahead of use (row-owner epilogues only; the coalesced residual class reads whole lines).  When the epilogue is the slower side (K = 1536 with an fp32 residual) it
        // starts a tile the moment it has finished the previous one, so a prefetch of THIS tile would be issued only
        // nanoseconds before its first loads; the tile this CTA drains NEXT is prefetched instead (one K loop ahead),
        // and the very first tile prefetches itself.
        auto prefetch_tile = [&](int t) {
          int mb, nb;
          tile_coords(t, num_m, num_n, group_m, mb, nb);
          const int r = mb * 2 * BM + (int)rank * BM + row_in_tile;
          const int c0 = nb * BN + col_half * 128;
          if (r < M && c0 < N) {
            const char* rp = reinterpret_cast<const char*>(ep.residual + (long long)r * ep.ldr + c0);
            const int cols = min(128, N - c0);
            for (int b = 0; b < cols * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + b));
          }
        };
        if (tile == pair) prefetch_tile(tile);
        if (tile + num_pairs < num_tiles) prefetch_tile(tile + num_pairs);
      }
      // per-column vectors of this tile -> shared memory while the K loop still runs (every lane needs the same 8
      // values at a time: an L1/L2 round trip per group of 8 columns was half of all epilogue stall samples)
      float* sbias = vec_smem + acc * 4 * BN;
      float* sgate = sbias + BN;
      float* slnu = sbias + 2 * BN;
      float* sgnext = sbias + 3 * BN;
      {
        const int t = threadIdx.x - 64;          // 0..255 = column inside the tile
        const int n = n_blk * BN + t;
        if (ep.bias) sbias[t] = n < N ? __ldg(ep.bias + n) : 0.f;
        if (ep.gate) sgate[t] = n < N ? __ldg(ep.gate + n) : 0.f;
        if (ep.ln_stats) slnu[t] = n < N ? __ldg(ep.ln_u + n) : 0.f;
        if (ep.a_next) sgnext[t] = n < N ? __ldg(ep.g_next + n) : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");   // epilogue warps only
      }
      mbar_wait_a(bar(TMEM_FULL + acc), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      float ss = 0.f;
      int ss_group = -1;
      float ln_r = 1.f, ln_nmu = 0.f;     // LN fold: rstd and -mean of this thread's row
      if (ep.ln_stats && row_ok) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(ep.ln_stats) + row);
        const float mean = st.x * ep.ln_inv_d;
        ln_r = rsqrtf(fmaxf(st.y * ep.ln_inv_d - mean * mean, 0.f) + ep.ln_eps);
        ln_nmu = -mean;
      }
      float ps = 0.f, pq = 0.f;           // producer: sum / sum of squares of the final values of this row (this tile half)
      // Residual-class epilogue (fp32 out = residual + gate * (acc + bias): o / cross-o / ffn.2 / patch embedding).  With the
      // thread = row mapping of tcgen05.ld every 16-byte access of a warp touches 32 different rows, and the LSU tag stage
      // (one 128-byte line per clock) capped the o-projection class at 2.2 TB/s (219 us against a 100 us HBM floor).  Here
      // the 32 x 32 chunk is transposed through a warp-private swizzled shared-memory tile so that global loads / stores
      // run with 8 lanes per 128-byte row segment: 4 lines per instruction instead of 32.
      const bool coal = ep.out_is_f32 && ep.residual && ep.act == SVI_ACT_NONE && !ep.sumsq && !ep.ln_stats;
      uint8_t* xt = xpose_smem + (warp - 2) * 4096;
      const int cq = lane & 7, cg = lane >> 3;          // column chunk (4 floats) / row group of this lane in the coalesced phase
      float ps8[8], pq8[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) ps8[it] = pq8[it] = 0.f;
#pragma unroll 1
      for (int c = col_half * 4; c < col_half * 4 + 4; ++c) {
        const int n0 = n_blk * BN + c * 32;
        if (n0 >= N) break;
        const bool whole = n0 + 32 <= N;     // all 32 columns of the chunk exist (always, except in the last N tile)
        if (coal && whole) {
          const int row_w0 = m_blk * 2 * BM + (int)rank * BM + quad * 32;      // first row of this warp's 32
          float4 res[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int gr = row_w0 + it * 4 + cg;
            res[it] = gr < M ? *reinterpret_cast<const float4*>(ep.residual + (long long)gr * ep.ldr + n0 + cq * 4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          uint32_t r[32];
          tmem_ld32(t_base + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q) {       // bias / gate in the row-owner form, then into the transposition tile
            float4 v = make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]), __uint_as_float(r[q * 4 + 2]),
                                   __uint_as_float(r[q * 4 + 3]));
            if (ep.bias) {
              const float4 b = *reinterpret_cast<const float4*>(sbias + c * 32 + q * 4);
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if (ep.gate) {
              const float4 g = *reinterpret_cast<const float4*>(sgate + c * 32 + q * 4);
              v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
            }
            *reinterpret_cast<float4*>(xt + lane * 128 + ((q ^ (lane & 7)) << 4)) = v;
          }
          __syncwarp();
          float4 gn = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ep.a_next) gn = *reinterpret_cast<const float4*>(sgnext + c * 32 + cq * 4);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + cg;
            const int gr = row_w0 + rr;
            float4 v = *reinterpret_cast<const float4*>(xt + rr * 128 + ((cq ^ (rr & 7)) << 4));
            v.x += res[it].x; v.y += res[it].y; v.z += res[it].z; v.w += res[it].w;
            if (gr < M) {
              if (ep.a_next) {
                ps8[it] += (v.x + v.y) + (v.z + v.w);
                pq8[it] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                uint2 an;
                an.x = pack_bf16x2(v.x * gn.x, v.y * gn.y);
                an.y = pack_bf16x2(v.z * gn.z, v.w * gn.w);
                *reinterpret_cast<uint2*>(ep.a_next + (long long)gr * ep.ld_an + n0 + cq * 4) = an;
              }
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (long long)gr * ep.ldo + n0 + cq * 4) = v;
            }
          }
          __syncwarp();       // the tile is rewritten by the next chunk
          continue;
        }
        // residual of the whole chunk first: eight independent 16-byte loads in flight under the TMEM load (with an
        // in-place residual the compiler may not move them above the stores of the previous group itself)
        float4 res[8];
        if (ep.residual && row_ok && whole) {
          const float4* rp = reinterpret_cast<const float4*>(ep.residual + (long long)row * ep.ldr + n0);
#pragma unroll
          for (int q = 0; q < 8; ++q) res[q] = rp[q];
        }
        uint32_t r[32];
        tmem_ld32(t_base + c * 32, r);
        tmem_ld_wait();
        if (row_ok) {
          if (ep.sumsq) sumsq_step(ep, n0, row, ss, ss_group);
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {
            const int n = n0 + j8 * 8;
            if (n >= N) break;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j8 * 8 + j]);
            if (ep.ln_stats) {
              const float4 u0 = *reinterpret_cast<const float4*>(slnu + c * 32 + j8 * 8);
              const float4 u1 = *reinterpret_cast<const float4*>(slnu + c * 32 + j8 * 8 + 4);
              v[0] = ln_r * fmaf(ln_nmu, u0.x, v[0]); v[1] = ln_r * fmaf(ln_nmu, u0.y, v[1]);
              v[2] = ln_r * fmaf(ln_nmu, u0.z, v[2]); v[3] = ln_r * fmaf(ln_nmu, u0.w, v[3]);
              v[4] = ln_r * fmaf(ln_nmu, u1.x, v[4]); v[5] = ln_r * fmaf(ln_nmu, u1.y, v[5]);
              v[6] = ln_r * fmaf(ln_nmu, u1.z, v[6]); v[7] = ln_r * fmaf(ln_nmu, u1.w, v[7]);
            }
            if (ep.bias) {
              const float4 b0 = *reinterpret_cast<const float4*>(sbias + c * 32 + j8 * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(sbias + c * 32 + j8 * 8 + 4);
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (ep.act != SVI_ACT_NONE) apply_act8(v, ep.act);
            if (ep.sumsq) {
#pragma unroll
              for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
            }
            if (ep.gate) {
              const float4 g0 = *reinterpret_cast<const float4*>(sgate + c * 32 + j8 * 8);
              const float4 g1 = *reinterpret_cast<const float4*>(sgate + c * 32 + j8 * 8 + 4);
              v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w;
              v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
            }
            if (ep.residual) {
              float4 r0, r1;
              if (whole) {
                r0 = res[2 * j8];
                r1 = res[2 * j8 + 1];
              } else {
                const float* rp = ep.residual + (long long)row * ep.ldr + n;
                r0 = *reinterpret_cast<const float4*>(rp);
                r1 = *reinterpret_cast<const float4*>(rp + 4);
              }
              v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
              v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
            }
            if (ep.a_next) {      // the next GEMM's LN-folded operand + the row statistics its epilogue needs
              const float4 g0 = *reinterpret_cast<const float4*>(sgnext + c * 32 + j8 * 8);
              const float4 g1 = *reinterpret_cast<const float4*>(sgnext + c * 32 + j8 * 8 + 4);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                ps += v[j];
                pq = fmaf(v[j], v[j], pq);
              }
              uint4 an;
              an.x = pack_bf16x2(v[0] * g0.x, v[1] * g0.y);
              an.y = pack_bf16x2(v[2] * g0.z, v[3] * g0.w);
              an.z = pack_bf16x2(v[4] * g1.x, v[5] * g1.y);
              an.w = pack_bf16x2(v[6] * g1.z, v[7] * g1.w);
              *reinterpret_cast<uint4*>(ep.a_next + (long long)row * ep.ld_an + n) = an;
            }
            if (ep.out_is_f32) {
              float* op = reinterpret_cast<float*>(ep.out) + (long long)row * ep.ldo + n;
              *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row * ep.ldo + n;
              uint4 pk;
              pk.x = pack_bf16x2(v[0], v[1]);
              pk.y = pack_bf16x2(v[2], v[3]);
              pk.z = pack_bf16x2(v[4], v[5]);
              pk.w = pack_bf16x2(v[6], v[7]);
              *reinterpret_cast<uint4*>(op) = pk;
            }
          }
        }
      }
      if (row_ok && ep.sumsq) sumsq_flush(ep, row, ss, ss_group);
      if (ep.a_next) {
        if (coal) {       // row sums of the coalesced phase: 8 lanes share a row
          const int row_w0 = m_blk * 2 * BM + (int)rank * BM + quad * 32;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            float a = ps8[it], b = pq8[it];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              a += __shfl_xor_sync(0xffffffffu, a, o);
              b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            const int gr = row_w0 + it * 4 + cg;
            if (cq == 0 && gr < M) {
              atomicAdd(&ep.row_stats[2LL * gr], a);
              atomicAdd(&ep.row_stats[2LL * gr + 1], b);
            }
          }
        }
        if (row_ok && (ps != 0.f || pq != 0.f)) {     // chunks that took the row-owner path (ragged last N tile)
          atomicAdd(&ep.row_stats[2LL * row], ps);
          atomicAdd(&ep.row_stats[2LL * row + 1], pq);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar(TMEM_EMPTY + acc), 0);   // the even CTA's barrier
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody frees TMEM / exits while the peer may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

int launch(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const svi_gemm_epilogue* e,
           cudaStream_t stream) {
  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, W, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, BK, BN_HALF);
  if (rc) return rc;
  Epi ep;
  ep.out = e->out; ep.ldo = e->ldo; ep.out_is_f32 = e->out_is_f32; ep.act = e->act;
  ep.bias = e->bias; ep.gate = e->gate; ep.residual = e->residual; ep.ldr = e->ldr;
  ep.sumsq = e->sumsq; ep.sumsq_groups = e->sumsq_groups; ep.sumsq_group_cols = e->sumsq_group_cols; ep.sumsq_parts = e->sumsq_parts;
  ep.ln_stats = e->ln_stats; ep.ln_u = e->ln_u; ep.ln_inv_d = e->ln_dim > 0 ? 1.0f / (float)e->ln_dim : 0.f; ep.ln_eps = e->ln_eps;
  ep.a_next = reinterpret_cast<__nv_bfloat16*>(e->a_next); ep.ld_an = e->ld_an; ep.g_next = e->g_next; ep.row_stats = e->row_stats;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(gemm2_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          smem_bytes(MAX_STAGES, false) > smem_bytes(MAX_STAGES - 1, true) ? smem_bytes(MAX_STAGES, false)
                                                                                                           : smem_bytes(MAX_STAGES - 1, true));
    if (ce != cudaSuccess) {
      set_last_error("svi_gemm_bf16(pair): cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce));
      return SVI_ERR_LAUNCH;
    }
    attr_set = true;
  }
  const int num_tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
  const int sms = sm_count();
  if (sms <= 0) return SVI_ERR_DRIVER;
  int pairs = sms / 2;
  if (num_tiles < pairs) pairs = num_tiles;
  // the residual-class epilogue (same predicate as `coal` in the kernel) trades one operand stage for its transposition tiles
  const bool xpose = ep.out_is_f32 && ep.residual && ep.act == SVI_ACT_NONE && !ep.sumsq && !ep.ln_stats;
  const int stages = xpose ? MAX_STAGES - 1 : MAX_STAGES;
  gemm2_bf16_kernel<<<2 * pairs, NUM_THREADS, smem_bytes(stages, xpose), stream>>>(ta, tb, M, N, K, GROUP_M, stages, ep);
  SVI_CUDA_LAUNCH_CHECK("svi_gemm_bf16(pair)");
  return SVI_OK;
}

}  // namespace gemm2
}  // namespace svi
