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
#include "common.cuh"
#include "../../include/svi_b200.h"

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

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + 2 * Q_TILE_BYTES;
  uint8_t* smem_v = smem_k + KV_STAGES * KV_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + KV_STAGES * KV_TILE_BYTES);
  uint64_t* q_full = bars;                    // [2]
  uint64_t* k_full = bars + 2;                // [4]
  uint64_t* k_empty = bars + 6;               // [4]
  uint64_t* v_full = bars + 10;               // [4]
  uint64_t* v_empty = bars + 14;              // [4]
  uint64_t* s_full = bars + 18;               // [2][2]  index i*2+b
  uint64_t* p_ready = bars + 22;              // [2][2]  index i*2+b (a softmax warpgroup may run ONE tile ahead of
                                              //         the MMA warp, so consecutive tiles must not share a barrier)
  uint64_t* pv_done = bars + 26;              // [2][2]  index i*2+b
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 30);

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
      for (int i = 0; i < 2; ++i) mbar_init(&q_full[i], 1);
      for (int i = 0; i < KV_STAGES; ++i) {
        mbar_init(&p_ready[i], 4);  // one arrive per softmax warp
        mbar_init(&k_full[i], 1);
        mbar_init(&k_empty[i], 1);
        mbar_init(&v_full[i], 1);
        mbar_init(&v_empty[i], 1);
        mbar_init(&s_full[i], 1);
        mbar_init(&pv_done[i], 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 8) {
    // ------------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      const int col0 = head * HD;
      mbar_expect_tx(&q_full[0], Q_TILE_BYTES);
      tma_load_2d(smem_q, &tmap_q, &q_full[0], col0, q_row0);
      tma_load_2d(smem_q + Q_HALF_BYTES, &tmap_q, &q_full[0], col0 + 64, q_row0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], KV_TILE_BYTES);
        tma_load_2d(smem_k + s * KV_TILE_BYTES, &tmap_k, &k_full[s], col0, j * BKV);
        tma_load_2d(smem_k + s * KV_TILE_BYTES + KV_HALF_BYTES, &tmap_k, &k_full[s], col0 + 64, j * BKV);
        if (j == 0) {
          mbar_expect_tx(&q_full[1], Q_TILE_BYTES);
          tma_load_2d(smem_q + Q_TILE_BYTES, &tmap_q, &q_full[1], col0, q_row0 + BQ);
          tma_load_2d(smem_q + Q_TILE_BYTES + Q_HALF_BYTES, &tmap_q, &q_full[1], col0 + 64, q_row0 + BQ);
        }
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_expect_tx(&v_full[s], KV_TILE_BYTES);
        tma_load_2d(smem_v + s * KV_TILE_BYTES, &tmap_v, &v_full[s], col0, j * BKV);
        tma_load_2d(smem_v + s * KV_TILE_BYTES + KV_HALF_BYTES, &tmap_v, &v_full[s], col0 + 64, j * BKV);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------ MMA issuer --------------------------------------
    constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);  // Q, K both K-major
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, HD, 0, 1);   // P K-major (TMEM), V MN-major
    constexpr uint32_t hi_kmaj = smem_desc_hi(1024, 2);
    const uint32_t lead = (lane == 0) ? 1u : 0u;
    const uint32_t q_lo = smem_desc_lo(smem_u32(smem_q), 16);
    const uint32_t k_lo = smem_desc_lo(smem_u32(smem_k), 16);
    const uint32_t v_lo = smem_desc_lo(smem_u32(smem_v), KV_HALF_BYTES);  // MN-major: LBO = distance of the 64-col atoms
    // whole warp executes (uniform operands -> uniform registers); `lead` predicates the single issuing lane
    auto issue_qk = [&](int i, int ks, int b) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const uint32_t offq = ((kk >> 2) * Q_HALF_BYTES + (kk & 3) * 32) >> 4;
        const uint32_t offk = ((kk >> 2) * KV_HALF_BYTES + (kk & 3) * 32) >> 4;
        tc_mma_ss_p(lead, tmem_base + i * 128 + b * 64, q_lo + ((i * Q_TILE_BYTES) >> 4) + offq, hi_kmaj,
                    k_lo + ((ks * KV_TILE_BYTES) >> 4) + offk, hi_kmaj, idesc_qk, kk != 0);
      }
    };
    auto issue_pv = [&](int i, int vs, int b, bool first_tile) {
#pragma unroll
      for (int kk = 0; kk < BKV / 16; ++kk)
        tc_mma_ts_p(lead, tmem_base + 256 + i * 128, tmem_base + i * 128 + b * 64 + kk * 8,
                    v_lo + ((vs * KV_TILE_BYTES + kk * 2048) >> 4), hi_kmaj, idesc_pv, !(first_tile && kk == 0));
    };

    mbar_wait(&k_full[0], 0);
    for (int i = 0; i < 2; ++i) {
      mbar_wait(&q_full[i], 0);
      tc_fence_after();
      issue_qk(i, 0, 0);
      tc_commit_p(lead, &s_full[i * 2 + 0]);
    }
    tc_commit_p(lead, &k_empty[0]);

    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) {  // next tile's S for both Q tiles goes first: it does not depend on this tile's softmax
        const int ks = (j + 1) % KV_STAGES;
        const int b = (j + 1) & 1;
        mbar_wait(&k_full[ks], ((j + 1) / KV_STAGES) & 1);
        tc_fence_after();
        issue_qk(0, ks, b);
        tc_commit_p(lead, &s_full[0 * 2 + b]);
        issue_qk(1, ks, b);
        tc_commit_p(lead, &s_full[1 * 2 + b]);
        tc_commit_p(lead, &k_empty[ks]);
      }
      const int vs = j % KV_STAGES;
      const int b = j & 1;
      mbar_wait(&v_full[vs], (j / KV_STAGES) & 1);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mbar_wait(&p_ready[i * 2 + b], (j >> 1) & 1);
        tc_fence_after();
        issue_pv(i, vs, b, j == 0);
        tc_commit_p(lead, &pv_done[i * 2 + b]);
      }
      tc_commit_p(lead, &v_empty[vs]);
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
      mbar_wait(&s_full[i * 2 + b], (j >> 1) & 1);
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
      float m8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) m8[u] = fmaxf(__uint_as_float(sr[0][u]), __uint_as_float(sr[0][u + 8]));
#pragma unroll
      for (int e = 16; e < 32; ++e) m8[e & 7] = fmaxf(m8[e & 7], __uint_as_float(sr[0][e]));
#pragma unroll
      for (int e = 0; e < 32; ++e) m8[e & 7] = fmaxf(m8[e & 7], __uint_as_float(sr[1][e]));
      float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      mx *= c;
      const bool need = (j > 0) && (mx > m_cur + RESCALE_THRESHOLD);
      if (j == 0) {
        m_cur = mx;
      } else if (__any_sync(0xffffffffu, need)) {
        // O_i must be quiescent: PV_i(j-1) has to retire first (S(j) can arrive before it in this pipeline)
        mbar_wait(&pv_done[i * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
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
          const float p0 = ex2(fmaf(__uint_as_float(sr[cc][e]), c, -m_cur));
          const float p1 = ex2(fmaf(__uint_as_float(sr[cc][e + 1]), c, -m_cur));
          l4[(e >> 1) & 3] += p0 + p1;
          pk[e >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st16(tS + cc * 16, pk);
      }
      l += (l4[0] + l4[1]) + (l4[2] + l4[3]);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[i * 2 + b]);
    }

    // epilogue: wait for the last PV, O / l -> bf16 -> global
    mbar_wait(&pv_done[i * 2 + ((n_kv - 1) & 1)], ((n_kv - 1) >> 1) & 1);
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
    cudaError_t ce = cudaFuncSetAttribute(attn2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
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
  attn2_fwd_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p);
  SVI_CUDA_LAUNCH_CHECK("svi_attn_fwd(v2)");
  return SVI_OK;
}

}  // namespace attn2
}  // namespace svi
