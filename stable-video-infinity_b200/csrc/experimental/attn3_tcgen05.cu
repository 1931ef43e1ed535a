// Flash attention forward v3 for sm_100a (head_dim 128, bf16 operands, fp32 softmax): v1's tile shape and MMA
// stream (two 128-row Q tiles per CTA, 128-row K/V tiles, S/P/O in TMEM, all MMAs at the full-rate N=128 shape) with
// FOUR softmax warpgroups instead of two: every 128x128 S tile is split column-wise between two warpgroups
// (thread = one row x 64 keys).  Measurements on v1/v2 (profiles/README.md) show the softmax side is latency bound,
// not pipe bound (XU 53 %, FMA/ALU far lower): 4 softmax warps per SM sub-partition instead of 2 hide the
// TMEM-load / MUFU / barrier latencies of one another and halve the per-tile critical path
//     S ready -> softmax -> P ready -> P*V -> Q*K^T(next) -> S ready.
// The two halves exchange their row maxima through shared memory (one named barrier per tile), hand their 64-key
// half of P to the MMA warp separately (P*V of half 0 overlaps the exponentials of half 1), and each rescales /
// stores its own 64 output columns.
//
//   warps  0-3  : Q tile 0, keys  0-63 of every K/V tile      warps  4-7  : Q tile 0, keys 64-127
//   warps  8-11 : Q tile 1, keys  0-63                         warps 12-15 : Q tile 1, keys 64-127
//   warp  16    : TMA producer          warp 17 : tcgen05.mma issuer + TMEM owner        warps 18-19 : idle
#include <stdlib.h>

#include "../common.cuh"
#include "../../../include/svi_b200.h"

namespace svi {
namespace attn3 {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int HD = 128;
constexpr int KV_STAGES = 2;
constexpr int HALF_BYTES = 128 * 64 * 2;    // one 128-row x 64-col swizzled box (16 KB)
constexpr int TILE_BYTES = 2 * HALF_BYTES;  // 128 x 128 bf16 (32 KB)
constexpr int NUM_THREADS = 640;
constexpr int TMEM_COLS = 512;
constexpr int XCHG_BYTES = 2 * 2 * 2 * 128 * 4;  // [parity][tile][half][row] floats
constexpr int SMEM_BYTES = (2 + 2 * KV_STAGES) * TILE_BYTES + XCHG_BYTES + 1024 + 512;
constexpr float RESCALE_THRESHOLD = 8.0f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 exp2_poly2(float2 x) {   // see attn_tcgen05.cu
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.f, 12582912.f));
  const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 q = __ffma2_rn(f, make_float2(0.055170949548482895f, 0.055170949548482895f),
                        make_float2(0.2426096349954605f, 0.2426096349954605f));
  q = __ffma2_rn(q, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  q = __ffma2_rn(q, f, make_float2(0.9999281764030457f, 0.9999281764030457f));
  q.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23));
  q.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23));
  return q;
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct Params {
  __nv_bfloat16* O;
  long long ldo;
  int Lq, Lk;
  float scale_log2;
  int accumulate;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
attn3_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  constexpr uint32_t OFF_Q = 0;
  constexpr uint32_t OFF_K = 2 * TILE_BYTES;
  constexpr uint32_t OFF_V = OFF_K + KV_STAGES * TILE_BYTES;
  constexpr uint32_t OFF_X = OFF_V + KV_STAGES * TILE_BYTES;   // row-max / row-sum exchange
  constexpr uint32_t OFF_BAR = OFF_X + XCHG_BYTES;
  enum : uint32_t { Q_FULL = 0, K_FULL = 2, K_EMPTY = 4, V_FULL = 6, V_EMPTY = 8, S_FULL = 10, P_READY = 12 /*[i*2+h]*/,
                    O_READY = 16 /*[i]*/, O_FULL = 18, NUM_BARS = 20 };
  auto bar = [&](uint32_t n) { return sbase + OFF_BAR + 8u * n; };
  uint8_t* const sgen = smem_raw + (sbase - sraw);   // generic pointer to the aligned base
  float* const xchg = reinterpret_cast<float*>(sgen + OFF_X);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sgen + OFF_BAR + 8 * NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q_row0 = blockIdx.x * (2 * BQ);
  const int n_kv = (p.Lk + BKV - 1) / BKV;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
  }
  if (warp == 17) {
    if (lane == 0) {
      for (uint32_t n = 0; n < NUM_BARS; ++n) {
        uint32_t count = 1;
        if (n >= P_READY && n < O_READY) count = 4;       // the 4 warps of one (tile, half) warpgroup
        if (n >= O_READY && n < O_FULL) count = 8;        // all 8 warps of a tile
        mbar_init_a(bar(n), count);
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
  if (*tmem_ptr_smem != 0u) {   // the CTA owns all 512 columns: the allocation must start at 0 (addresses are constants)
    if (threadIdx.x == 0) printf("svi: unexpected TMEM base 0x%x\n", *tmem_ptr_smem);
    __trap();
  }
  constexpr uint32_t tmem_base = 0;

  if (warp == 16) {
    // ------------------------------------ TMA producer ------------------------------------
    if (lane == 0) {
      const int col0 = head * HD;
      auto load_tile = [&](uint32_t dst, const CUtensorMap* m, uint32_t b, int row) {
        mbar_expect_tx_a(b, TILE_BYTES);
        tma_load_2d_a(dst, m, b, col0, row);
        tma_load_2d_a(dst + HALF_BYTES, m, b, col0 + 64, row);
      };
      load_tile(sbase + OFF_Q, &tmap_q, bar(Q_FULL + 0), q_row0);
      for (int j0 = 0; j0 < n_kv; j0 += KV_STAGES) {
        const uint32_t ph = (j0 / KV_STAGES) & 1;
#pragma unroll
        for (int s = 0; s < KV_STAGES; ++s) {
          const int j = j0 + s;
          if (j >= n_kv) break;
          mbar_wait_a(bar(K_EMPTY + s), ph ^ 1);
          load_tile(sbase + OFF_K + s * TILE_BYTES, &tmap_k, bar(K_FULL + s), j * BKV);
          if (j == 0) load_tile(sbase + OFF_Q + TILE_BYTES, &tmap_q, bar(Q_FULL + 1), q_row0 + BQ);
          mbar_wait_a(bar(V_EMPTY + s), ph ^ 1);
          load_tile(sbase + OFF_V + s * TILE_BYTES, &tmap_v, bar(V_FULL + s), j * BKV);
        }
      }
    }
  } else if (warp == 17) {
    // ------------------------------------ MMA issuer (whole warp, uniform operands) -------
    constexpr uint32_t idesc_qk = make_idesc_bf16(BQ, BKV, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(BQ, HD, 0, 1);
    constexpr uint32_t hi_kmaj = smem_desc_hi(1024, 2);
    const uint32_t q_lo = smem_desc_lo(sbase + OFF_Q, 16);
    const uint32_t k_lo = smem_desc_lo(sbase + OFF_K, 16);
    const uint32_t v_lo = smem_desc_lo(sbase + OFF_V, HALF_BYTES);
    auto issue_qk = [&](int i, int ks) {
      const uint32_t a0 = q_lo + ((i * TILE_BYTES) >> 4), b0 = k_lo + ((ks * TILE_BYTES) >> 4);
      tc_mma_ss_k4(tmem_base + i * 128, a0, hi_kmaj, b0, hi_kmaj, idesc_qk, 0);
      tc_mma_ss_k4(tmem_base + i * 128, a0 + (HALF_BYTES >> 4), hi_kmaj, b0 + (HALF_BYTES >> 4), hi_kmaj, idesc_qk, 1);
    };
    auto issue_pv_half = [&](int i, int vs, int h, uint32_t accumulate) {
      tc_mma_ts_k4(tmem_base + 256 + i * 128, tmem_base + i * 128 + h * 64, v_lo + ((vs * TILE_BYTES + h * 4 * 2048) >> 4),
                   hi_kmaj, idesc_pv, accumulate);
    };

    mbar_wait_a(bar(K_FULL + 0), 0);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_wait_a(bar(Q_FULL + i), 0);
      tc_fence_after();
      issue_qk(i, 0);
      tc_commit_a(bar(S_FULL + i));
    }
    tc_commit_a(bar(K_EMPTY + 0));

    for (int j0 = 0; j0 < n_kv; j0 += KV_STAGES) {
      const uint32_t ph = (j0 / KV_STAGES) & 1;
#pragma unroll
      for (int u = 0; u < KV_STAGES; ++u) {
        const int j = j0 + u;
        if (j >= n_kv) break;
        const bool has_next = (j + 1) < n_kv;
        const int ks = (u + 1) & 1;
        mbar_wait_a(bar(V_FULL + u), ph);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          mbar_wait_a(bar(O_READY + i), u & 1);          // both halves are done touching O_i (lazy rescale)
          mbar_wait_a(bar(P_READY + i * 2 + 0), u & 1);  // tile j: phase j & 1 = u & 1 (KV_STAGES == 2)
          tc_fence_after();
          issue_pv_half(i, u, 0, j > 0);
          if (has_next && i == 0) mbar_wait_a(bar(K_FULL + ks), (u == 1) ? (ph ^ 1) : ph);
          mbar_wait_a(bar(P_READY + i * 2 + 1), u & 1);
          tc_fence_after();
          issue_pv_half(i, u, 1, 1);
          if (has_next) {
            issue_qk(i, ks);
            tc_commit_a(bar(S_FULL + i));
          } else {
            tc_commit_a(bar(O_FULL + i));
          }
        }
        tc_commit_a(bar(V_EMPTY + u));
        if (has_next) tc_commit_a(bar(K_EMPTY + ks));
      }
    }
  } else if (warp < 16) {
    // ------------------------------------ softmax warpgroups ------------------------------
    const int i = warp >> 3;          // Q tile
    const int h = (warp >> 2) & 1;    // key half of every K/V tile
    const int quad = warp & 3;        // TMEM lane quadrant
    const int r = quad * 32 + lane;   // row inside the Q tile
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + i * 128 + h * 64 + lane_sel;    // my 64 S columns
    const uint32_t tP = tS;   // my 32 packed P columns alias the FIRST half of my own S columns only: the partner
                              // warpgroup may still be re-reading its S columns when I store P (two-pass softmax)
    const uint32_t tO = tmem_base + 256 + i * 128 + h * 64 + lane_sel;   // my 64 O columns
    const float c = p.scale_log2;
    float m_cur = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait_a(bar(S_FULL + i), j & 1);
      tc_fence_after();
      const int limit = p.Lk - j * BKV - h * 64;   // valid columns of my half (>= 64: all)
      float mx;
      {   // pass 1: row max of my 64 columns (registers die here: the launch cap is 96/thread with 20 warps)
        uint32_t sr[2][32];
        tmem_ld32(tS + 0, sr[0]);
        tmem_ld32(tS + 32, sr[1]);
        tmem_ld_wait();
        if (limit < 64) {
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
        mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      }
      // row max of the whole 128-key tile: exchange with the partner thread of the other half
      float* xb = xchg + (((j & 1) * 2 + i) * 2) * 128;
      xb[h * 128 + r] = mx;
      named_bar_sync(1 + i, 256);
      mx = fmaxf(mx, xb[(h ^ 1) * 128 + r]) * c;   // at least one valid key per tile -> finite
      // lazy rescale: both partner threads see the same (mx, m_cur) and take the same decision
      const bool need = (j > 0) && (mx > m_cur + RESCALE_THRESHOLD);
      if (j == 0) {
        m_cur = mx;
      } else if (__any_sync(0xffffffffu, need)) {
        const float m_new = fmaxf(m_cur, mx);
        const float alpha = ex2(m_cur - m_new);
        l *= alpha;
        m_cur = m_new;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t o[32];
          tmem_ld32(tO + cc * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
          tmem_st32(tO + cc * 32, o);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar(O_READY + i));   // my O columns are final for this tile's P*V
      // P = exp2(S*c - m): 5/8 on MUFU, 3/8 as FFMA2 polynomial; packed row-sum chains
      float2 l4[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_cur, -m_cur);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {   // pass 2: reload 32 columns from TMEM (the other warps of the sub-partition hide it)
        uint32_t sc[32];
        tmem_ld32(tS + cc * 32, sc);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const int pr = e >> 1;
          float2 x = __ffma2_rn(make_float2(__uint_as_float(sc[e]), __uint_as_float(sc[e + 1])), c2, nm2);
          if (limit < 64) {
            if (cc * 32 + e >= limit) x.x = -INFINITY;
            if (cc * 32 + e + 1 >= limit) x.y = -INFINITY;
          }
          float2 pv;
          if ((pr & 7) < 3) {
            pv = exp2_poly2(x);
          } else {
            pv.x = ex2(x.x);
            pv.y = ex2(x.y);
          }
          l4[pr & 3] = __fadd2_rn(l4[pr & 3], pv);
          pk[pr] = pack_bf16x2(pv.x, pv.y);
        }
        tmem_st16(tP + cc * 16, pk);
      }
      {
        const float2 a = __fadd2_rn(__fadd2_rn(l4[0], l4[1]), __fadd2_rn(l4[2], l4[3]));
        l += a.x + a.y;
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(bar(P_READY + i * 2 + h));
    }

    // epilogue: total row sum = both halves; O / l -> bf16 -> global (my 64 columns)
    float* xb = xchg + ((n_kv & 1) * 2 + i) * 2 * 128;
    xb[h * 128 + r] = l;
    named_bar_sync(1 + i, 256);
    const float inv_l = 1.0f / (l + xb[(h ^ 1) * 128 + r]);
    mbar_wait_a(bar(O_FULL + i), 0);
    tc_fence_after();
    const int row = q_row0 + i * BQ + r;
    __nv_bfloat16* orow = p.O + (long long)row * p.ldo + head * HD + h * 64;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t o[32];
      tmem_ld32(tO + cc * 32, o);
      tmem_ld_wait();
      if (row < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[g * 8 + e]) * inv_l;
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
  if (warp == 17) {
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
    cudaError_t ce = cudaFuncSetAttribute(attn3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce != cudaSuccess) {
      set_last_error("svi_attn_fwd(v3): cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce));
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
  attn3_fwd_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tk, tv, p);
  SVI_CUDA_LAUNCH_CHECK("svi_attn_fwd(v3)");
  return SVI_OK;
}

}  // namespace attn3
}  // namespace svi
