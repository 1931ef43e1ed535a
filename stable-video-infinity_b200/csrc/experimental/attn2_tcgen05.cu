// Flash attention forward v2 for sm_100a (head_dim 128, bf16 operands, fp32 softmax): software-pipelined so that
// the tensor pipe never waits for the softmax of the SAME tile.
//
// v1 (attn_tcgen05.cu) keeps one S buffer per Q tile, so per Q tile the chain  S(j) -> softmax -> P(j) -> PV(j)
// -> QK(j+1) -> S(j+1)  is serial and the tensor pipe idles for one softmax latency per K/V tile (ncu: tensor
// pipe 43 %, XU 44 %, nothing saturated).  v2 halves the K/V tile to 64 rows, which lets every Q tile own TWO
// S buffers inside the same 512 TMEM columns:
//
//   TMEM: S0[0] S0[1] S1[0] S1[1] (64 cols each) | O0 (128) | O1 (128);  P_i[b] (bf16) aliases S_i[b][0:32)
//
// QK_i(j+1) is issued into the other buffer BEFORE the softmax of tile j has finished, so S(j+1) is already
// waiting when the softmax warpgroup comes back, and PV_i(j) only waits for P_i(j).  Both softmax warpgroups
// and the tensor pipe now run concurrently; the steady state is bound by max(MMA, MUFU) instead of their sum.
// O rescaling stays lazy (2^8 threshold); because S(j) now arrives before PV(j-1) has retired, a rescale first
// waits on pv_done[i][(j-1)&1] (phase-exact: that barrier can be at most one completion ahead).
//
//   warps 0-3 / 4-7 : softmax warpgroup of Q tile 0 / 1 (thread = row)
//   warp 8          : TMA producer (Q once; K_j, V_j 64-row tiles into 4-stage rings)
//   warp 9          : tcgen05.mma issuer + TMEM owner
#include <stdlib.h>

#include "../common.cuh"
#include "../../../include/svi_b200.h"

namespace svi {
namespace attn2 {

constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int HD = 128;
constexpr int KV_STAGES = 4;
static_assert(KV_STAGES == 4, "barrier initialisation below assumes 4 K/V stages (= 2x2 S / PV barriers)");
constexpr int Q_HALF_BYTES = 128 * 64 * 2;    // 16 KB: 128 rows x 64 cols, 128B-swizzled
constexpr int Q_TILE_BYTES = 2 * Q_HALF_BYTES;
constexpr int KV_HALF_BYTES = 64 * 64 * 2;    // 8 KB: 64 rows x 64 cols
constexpr int KV_TILE_BYTES = 2 * KV_HALF_BYTES;
constexpr int NUM_THREADS = 320;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = 2 * Q_TILE_BYTES + 2 * KV_STAGES * KV_TILE_BYTES + 1024 + 512;
constexpr float RESCALE_THRESHOLD = 8.0f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Params {
  __nv_bfloat16* O;
  long long ldo;
  int Lq, Lk;
  float scale_log2;
  int accumulate;
};

// DBG (development only, SVI_ATTN_DEBUG env): 0 = product kernel; 1 = exponentials replaced by FMAs (no MUFU);
// 2 = softmax warps only move S -> P (no max / exp / sum); 3 = 2 + the MMA warp ignores P_READY; 4 = 3 + the MMA warp
// ignores K_FULL / V_FULL (results are garbage for DBG >= 1) — timing probes that isolate the MMA / load side.
template <int DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, Params p) {
  extern __shared__ uint8_t smem_raw[];
  // everything below works on 32-bit shared-window addresses (uniform values -> uniform registers)
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  constexpr uint32_t OFF_Q = 0;
  constexpr uint32_t OFF_K = 2 * Q_TILE_BYTES;
  constexpr uint32_t OFF_V = OFF_K + KV_STAGES * KV_TILE_BYTES;
  constexpr uint32_t OFF_BAR = OFF_V + KV_STAGES * KV_TILE_BYTES;
  enum : uint32_t { Q_FULL = 0, K_FULL = 2, K_EMPTY = 6, V_FULL = 10, V_EMPTY = 14, S_FULL = 18, P_READY = 22, PV_DONE = 26,
                    NUM_BARS = 30 };
  // S_FULL / P_READY / PV_DONE are indexed [i*2+b]: a softmax warpgroup may run ONE tile ahead of the MMA warp, so
  // consecutive tiles must not share a barrier
  auto bar = [&](uint32_t n) { return sbase + OFF_BAR + 8u * n; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_raw + (sbase - sraw) + OFF_BAR + 8 * NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q_row0 = blockIdx.x * (2 * BQ);
  const int n_kv = (p.Lk + BKV - 1) / BKV;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 9) {
    if (lane == 0) {
      for (uint32_t n = 0; n < NUM_BARS; ++n) mbar_init_a(bar(n), (n >= P_READY && n < PV_DONE) ? 4 : 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // The CTA owns all 512 TMEM columns (1 CTA / SM), so the allocation starts at column 0 / lane 0: TMEM addresses
  // below are compile-time constants.  Verified here; anything else is a hard error.
  if (*tmem_ptr_smem != 0u) {
    if (threadIdx.x == 0) printf("svi: unexpected TMEM base 0x%x\n", *tmem_ptr_smem);
    __trap();
  }
  constexpr uint32_t tmem_base = 0;

  if (warp == 8) {
    // ------------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      const int col0 = head * HD;
      mbar_expect_tx_a(bar(Q_FULL + 0), Q_TILE_BYTES);
      tma_load_2d_a(sbase + OFF_Q, &tmap_q, bar(Q_FULL + 0), col0, q_row0);
      tma_load_2d_a(sbase + OFF_Q + Q_HALF_BYTES, &tmap_q, bar(Q_FULL + 0), col0 + 64, q_row0);
      for (int j0 = 0; j0 < n_kv; j0 += KV_STAGES) {
        const uint32_t ph = (j0 / KV_STAGES) & 1;
#pragma unroll
        for (int s = 0; s < KV_STAGES; ++s) {
          const int j = j0 + s;
          if (j >= n_kv) break;
          mbar_wait_a(bar(K_EMPTY + s), ph ^ 1);
          mbar_expect_tx_a(bar(K_FULL + s), KV_TILE_BYTES);
          tma_load_2d_a(sbase + OFF_K + s * KV_TILE_BYTES, &tmap_k, bar(K_FULL + s), col0, j * BKV);
          tma_load_2d_a(sbase + OFF_K + s * KV_TILE_BYTES + KV_HALF_BYTES, &tmap_k, bar(K_FULL + s), col0 + 64, j * BKV);
          if (j == 0) {
            mbar_expect_tx_a(bar(Q_FULL + 1), Q_TILE_BYTES);
            tma_load_2d_a(sbase + OFF_Q + Q_TILE_BYTES, &tmap_q, bar(Q_FULL + 1), col0, q_row0 + BQ);
            tma_load_2d_a(sbase + OFF_Q + Q_TILE_BYTES + Q_HALF_BYTES, &tmap_q, bar(Q_FULL + 1), col0 + 64, q_row0 + BQ);
          }
          mbar_wait_a(bar(V_EMPTY + s), ph ^ 1);
          mbar_expect_tx_a(bar(V_FULL + s), KV_TILE_BYTES);
          tma_load_2d_a(sbase + OFF_V + s * KV_TILE_BYTES, &tmap_v, bar(V_FULL + s), col0, j * BKV);
          tma_load_2d_a(sbase + OFF_V + s * KV_TILE_BYTES + KV_HALF_BYTES, &tmap_v, bar(V_FULL + s), col0 + 64, j * BKV);
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------ MMA issuer --------------------------------------
    // The WHOLE warp runs this code with warp-uniform operands; the issuing lane is picked by elect.sync inside the
    // batched wrappers (4 K-steps per asm block).  Stage / buffer indices are compile-time constants (x4 unroll).
    constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, HD, 0, 1);   // P K-major (TMEM), V MN-major
    constexpr uint32_t hi_kmaj = smem_desc_hi(1024, 2);            // SBO 1024 B (8 rows x 128 B), 128B swizzle
    const uint32_t q_lo = smem_desc_lo(sbase + OFF_Q, 16);
    const uint32_t k_lo = smem_desc_lo(sbase + OFF_K, 16);
    const uint32_t v_lo = smem_desc_lo(sbase + OFF_V, KV_HALF_BYTES);  // MN-major: LBO = distance of the 64-col atoms
    auto issue_qk = [&](int i, int ks, int b) {   // S_i[b] = Q_i K^T : 8 K-steps = 2 swizzle boxes x 4
      const uint32_t d = tmem_base + i * 128 + b * 64;
      const uint32_t a0 = q_lo + ((i * Q_TILE_BYTES) >> 4), b0 = k_lo + ((ks * KV_TILE_BYTES) >> 4);
      tc_mma_ss_k4(d, a0, hi_kmaj, b0, hi_kmaj, idesc_qk, 0);
      tc_mma_ss_k4(d, a0 + (Q_HALF_BYTES >> 4), hi_kmaj, b0 + (KV_HALF_BYTES >> 4), hi_kmaj, idesc_qk, 1);
    };
    auto issue_pv = [&](int i, int vs, int b, uint32_t accumulate) {   // O_i (+)= P_i[b] V : 4 K-steps of 16 rows
      tc_mma_ts_k4(tmem_base + 256 + i * 128, tmem_base + i * 128 + b * 64, v_lo + ((vs * KV_TILE_BYTES) >> 4), hi_kmaj,
                   idesc_pv, accumulate);
    };

    mbar_wait_a(bar(K_FULL + 0), 0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_wait_a(bar(Q_FULL + i), 0);
      tc_fence_after();
      issue_qk(i, 0, 0);
      tc_commit_a(bar(S_FULL + i * 2 + 0));
    }
    tc_commit_a(bar(K_EMPTY + 0));

    for (int j0 = 0; j0 < n_kv; j0 += KV_STAGES) {
      const uint32_t ph = (j0 / KV_STAGES) & 1;
#pragma unroll
      for (int u = 0; u < KV_STAGES; ++u) {
        const int j = j0 + u;
        if (j >= n_kv) break;
        if (j + 1 < n_kv) {  // next tile's S for both Q tiles goes first: it does not depend on this tile's softmax
          constexpr int dummy = 0;
          (void)dummy;
          const int ks = (u + 1) & 3;
          const int b1 = (u + 1) & 1;
          if (DBG < 4) mbar_wait_a(bar(K_FULL + ks), (u == 3) ? (ph ^ 1) : ph);
          tc_fence_after();
          issue_qk(0, ks, b1);
          tc_commit_a(bar(S_FULL + 0 * 2 + b1));
          issue_qk(1, ks, b1);
          tc_commit_a(bar(S_FULL + 1 * 2 + b1));
          tc_commit_a(bar(K_EMPTY + ks));
        }
        const int b = u & 1;
        if (DBG < 4) mbar_wait_a(bar(V_FULL + u), ph);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (DBG < 3) mbar_wait_a(bar(P_READY + i * 2 + b), (u >> 1) & 1);   // tile j uses phase (j >> 1) & 1 = (u >> 1) & 1
          tc_fence_after();
          issue_pv(i, u, b, j != 0);
          tc_commit_a(bar(PV_DONE + i * 2 + b));
        }
        tc_commit_a(bar(V_EMPTY + u));
      }
    }
  } else {
    // ------------------------------------ softmax warpgroups ------------------------------
    const int i = warp >> 2;
    const int quad = warp & 3;
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tO = tmem_base + 256 + i * 128 + lane_sel;
    const float c = p.scale_log2;
    float m_cur = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const uint32_t tS = tmem_base + i * 128 + b * 64 + lane_sel;
      mbar_wait_a(bar(S_FULL + i * 2 + b), (j >> 1) & 1);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld32(tS + 0, sr[0]);
      tmem_ld32(tS + 32, sr[1]);
      tmem_ld_wait();
      const int limit = p.Lk - j * BKV;
      if (limit < BKV) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (cc * 32 + e >= limit) sr[cc][e] = 0xff800000u;  // -inf
      }
      if (DBG >= 2) m_cur = 0.f;
      float m8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) m8[u] = fmaxf(__uint_as_float(sr[0][u]), __uint_as_float(sr[0][u + 8]));
#pragma unroll
      for (int e = 16; e < 32; ++e) m8[e & 7] = fmaxf(m8[e & 7], __uint_as_float(sr[0][e]));
#pragma unroll
      for (int e = 0; e < 32; ++e) m8[e & 7] = fmaxf(m8[e & 7], __uint_as_float(sr[1][e]));
      float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      mx *= c;
      const bool need = (DBG < 2) && (j > 0) && (mx > m_cur + RESCALE_THRESHOLD);
      if (j == 0) {
        if (DBG < 2) m_cur = mx;
      } else if (__any_sync(0xffffffffu, need)) {
        // O_i must be quiescent: PV_i(j-1) has to retire first (S(j) can arrive before it in this pipeline)
        mbar_wait_a(bar(PV_DONE + i * 2 + ((j - 1) & 1)), ((j - 1) >> 1) & 1);
        tc_fence_after();
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
      float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = fmaf(__uint_as_float(sr[cc][e]), c, -m_cur);
          float p1 = fmaf(__uint_as_float(sr[cc][e + 1]), c, -m_cur);
          if (DBG == 0) {
            p0 = ex2(p0);
            p1 = ex2(p1);
          }
          l4[(e >> 1) & 3] += p0 + p1;
          pk[e >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st16(tS + cc * 16, pk);
      }
      l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar(P_READY + i * 2 + b));
    }

    // epilogue: wait for the last PV, O / l -> bf16 -> global
    mbar_wait_a(bar(PV_DONE + i * 2 + ((n_kv - 1) & 1)), ((n_kv - 1) >> 1) & 1);
    tc_fence_after();
    const int row = q_row0 + i * BQ + quad * 32 + lane;
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = p.O + (long long)row * p.ldo + head * HD;
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

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int launch(const void* Q, long long ldq, const void* K, long long ldk, const void* V, long long ldv, void* O,
           long long ldo, int Lq, int Lk, int num_heads, float scale, int accumulate, cudaStream_t stream) {
  const int64_t width = (int64_t)num_heads * HD;
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_2d(&tq, Q, 2, (uint64_t)width, (uint64_t)Lq, (uint64_t)ldq * 2, 64, BQ);
  if (rc) return rc;
  rc = make_tmap_2d(&tk, K, 2, (uint64_t)width, (uint64_t)Lk, (uint64_t)ldk * 2, 64, BKV);
  if (rc) return rc;
  rc = make_tmap_2d(&tv, V, 2, (uint64_t)width, (uint64_t)Lk, (uint64_t)ldv * 2, 64, BKV);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(attn2_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(attn2_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(attn2_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(attn2_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(attn2_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce != cudaSuccess) {
      set_last_error("svi_attn_fwd(v2): cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce));
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
  dim3 grid((Lq + 2 * BQ - 1) / (2 * BQ), num_heads);
  static const int dbg = []() { const char* e = getenv("SVI_ATTN_DEBUG"); return e ? atoi(e) : 0; }();
  if (dbg == 1) attn2_fwd_kernel<1><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p);
  else if (dbg == 2) attn2_fwd_kernel<2><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p);
  else if (dbg == 3) attn2_fwd_kernel<3><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p);
  else if (dbg == 4) attn2_fwd_kernel<4><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p);
  else attn2_fwd_kernel<0><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p);
  SVI_CUDA_LAUNCH_CHECK("svi_attn_fwd(v2)");
  return SVI_OK;
}

}  // namespace attn2
}  // namespace svi
