// Persistent warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A[M,K] @ W[N,K]^T).
//
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B-swizzled 128x64 / 256x64 operand tiles)
//   warp 1      : tcgen05.mma issuer (one lane), owns the TMEM allocation (512 columns = 2 accumulators)
//   warps 2..5  : epilogue (tcgen05.ld -> bias / activation / sum-of-squares / gate / residual -> global)
//
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, double buffered so
// the epilogue of tile i overlaps the MMAs of tile i+1), and the static persistent tile loop.
// Replaces every F.linear of the reference's DiT (see include/svi_b200.h for the line map).
#include <cstdlib>
#include "common.cuh"
#include "../../include/svi_b200.h"
#include "gemm_epilogue.cuh"

namespace svi {
namespace gemm {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

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
};


__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a,
                 const __grid_constant__ CUtensorMap tmap_b, int M, int N, int K, Epi ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full_bar = bars + 2 * STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full_bar[i], 1);
        mbar_init(&tmem_empty_bar[i], 4);  // one arrive per epilogue warp
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

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d(sa + A_STAGE_BYTES, &tmap_b, &full_bar[stage], kb * BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    constexpr uint32_t hi_kmaj = smem_desc_hi(1024, 2);  // SBO 1024 B, 128B swizzle, K-major
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        {
          // whole warp executes with uniform operands; elect.sync inside the wrappers picks the issuing lane
          const uint32_t a_lo = smem_desc_lo(smem_u32(smem + stage * STAGE_BYTES), 16);
          const uint32_t b_lo = a_lo + (A_STAGE_BYTES >> 4);
          tc_mma_ss_k4(d_tmem, a_lo, hi_kmaj, b_lo, hi_kmaj, idesc, kb != 0);   // BK = 64 = 4 K-steps, one asm block
          tc_commit_p(0, &empty_bar[stage]);                       // smem slot free when MMAs retire
          if (kb == num_k - 1) tc_commit_p(0, &tmem_full_bar[acc]);  // accumulator ready
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ------------------------------- epilogue -----------------------------------
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int row = m_blk * BM + row_in_tile;
      const bool row_ok = row < M;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      float ss = 0.f;
      int ss_group = -1;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = n_blk * BN + c * 32;
        if (n0 >= N) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(t_base + c * 32, r);
        tmem_ld_wait();
        if (row_ok) {
          if (ep.sumsq) sumsq_step(ep, n0, row, ss, ss_group);
#pragma unroll
          for (int j8 = 0; j8 < 4; ++j8) {  // 8 columns at a time
            const int n = n0 + j8 * 8;
            if (n >= N) break;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j8 * 8 + j]);
            if (ep.bias) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + 4));
              v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
              v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
            }
            if (ep.act != SVI_ACT_NONE) apply_act8(v, ep.act);
            if (ep.sumsq) {
#pragma unroll
              for (int j = 0; j < 8; ++j) ss += v[j] * v[j];
            }
            if (ep.gate) {
              const float4 g0 = __ldg(reinterpret_cast<const float4*>(ep.gate + n));
              const float4 g1 = __ldg(reinterpret_cast<const float4*>(ep.gate + n + 4));
              v[0] *= g0.x; v[1] *= g0.y; v[2] *= g0.z; v[3] *= g0.w;
              v[4] *= g1.x; v[5] *= g1.y; v[6] *= g1.z; v[7] *= g1.w;
            }
            if (ep.residual) {
              const float* rp = ep.residual + (long long)row * ep.ldr + n;
              const float4 r0 = *reinterpret_cast<const float4*>(rp);
              const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
              v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
              v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
            }
            if (ep.out_is_f32) {
              float* op = reinterpret_cast<float*>(ep.out) + (long long)row * ep.ldo + n;
              *reinterpret_cast<float4*>(op) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(op + 4) = make_float4(v[4], v[5], v[6], v[7]);
            } else {
              __nv_bfloat16* op =
                  reinterpret_cast<__nv_bfloat16*>(ep.out) + (long long)row * ep.ldo + n;
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
      // all tcgen05.ld of this warp have completed (wait::ld above): hand the accumulator back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace gemm
}  // namespace svi

namespace svi { namespace gemm2 {
int launch(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K, const svi_gemm_epilogue* e,
           cudaStream_t stream);
} }

extern "C" int svi_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int32_t M,
                             int32_t N, int32_t K, const svi_gemm_epilogue* e, void* stream) {
  using namespace svi;
  using namespace svi::gemm;
  SVI_REQUIRE(A && W && e && e->out, "svi_gemm_bf16: null pointer");
  SVI_REQUIRE(M > 0 && N > 0 && K > 0, "svi_gemm_bf16: M,N,K must be positive (got %d,%d,%d)", M, N, K);
  SVI_REQUIRE(K % 8 == 0 && N % 8 == 0, "svi_gemm_bf16: K and N must be multiples of 8 (K=%d N=%d)", K, N);
  SVI_REQUIRE(lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K,
              "svi_gemm_bf16: lda/ldw must be multiples of 8 and >= K");
  SVI_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
              "svi_gemm_bf16: A and W must be 16-byte aligned");
  SVI_REQUIRE(e->ldo >= N && e->ldo % (e->out_is_f32 ? 4 : 8) == 0 &&
                  (reinterpret_cast<uintptr_t>(e->out) & 15) == 0,
              "svi_gemm_bf16: out must be 16-byte aligned with ldo >= N and 16-byte row pitch");
  SVI_REQUIRE(e->act >= 0 && e->act <= 4, "svi_gemm_bf16: unknown activation %d", e->act);
  if (e->residual)
    SVI_REQUIRE(e->ldr >= N && e->ldr % 4 == 0 && (reinterpret_cast<uintptr_t>(e->residual) & 15) == 0,
                "svi_gemm_bf16: residual must be 16-byte aligned with ldr >= N, ldr %% 4 == 0");
  if (e->bias) SVI_REQUIRE((reinterpret_cast<uintptr_t>(e->bias) & 15) == 0, "svi_gemm_bf16: bias alignment");
  if (e->gate) SVI_REQUIRE((reinterpret_cast<uintptr_t>(e->gate) & 15) == 0, "svi_gemm_bf16: gate alignment");
  if (e->sumsq)
    SVI_REQUIRE(e->sumsq_groups > 0 && e->sumsq_group_cols > 0 && e->sumsq_group_cols % 32 == 0,
                "svi_gemm_bf16: sumsq_group_cols must be a positive multiple of 32");
  if (e->sumsq && e->sumsq_parts)
    SVI_REQUIRE(e->sumsq_group_cols % 128 == 0 && e->sumsq_parts == e->sumsq_group_cols / 128,
                "svi_gemm_bf16: sumsq_parts must be 0 (atomic accumulation) or sumsq_group_cols / 128 with sumsq_group_cols %% 128 == 0");

  if (e->ln_stats || e->a_next) {
    SVI_REQUIRE(M > BM, "svi_gemm_bf16: the LayerNorm fold (ln_stats / a_next) runs in the CTA-pair kernel: needs M > 128");
    if (e->ln_stats)
      SVI_REQUIRE(e->ln_u && e->ln_dim > 0 && (reinterpret_cast<uintptr_t>(e->ln_stats) & 7) == 0,
                  "svi_gemm_bf16: ln fold needs ln_u, ln_dim > 0 and 8-byte aligned ln_stats");
    if (e->a_next)
      SVI_REQUIRE(e->g_next && e->row_stats && e->ld_an >= N && e->ld_an % 8 == 0 &&
                      (reinterpret_cast<uintptr_t>(e->a_next) & 15) == 0,
                  "svi_gemm_bf16: a_next needs g_next, row_stats, ld_an >= N (multiple of 8) and 16-byte alignment");
  }

  // More than one 128-row tile of A: CTA pairs (256x256 tiles, gemm2_tcgen05.cu; 12-13 % faster on the DiT shapes
  // because each SM pulls a third less operand data through L2).  M <= 128 (time/text embeddings, LoRA merges of
  // narrow matrices) would leave the odd CTA of every pair idle, so those stay on the single-CTA kernel below.
  if (M > BM) return svi::gemm2::launch(A, lda, W, ldw, M, N, K, e, static_cast<cudaStream_t>(stream));

  CUtensorMap ta, tb;
  int rc = make_tmap_2d(&ta, A, 2, (uint64_t)K, (uint64_t)M, (uint64_t)lda * 2, BK, BM);
  if (rc) return rc;
  rc = make_tmap_2d(&tb, W, 2, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, BK, BN);
  if (rc) return rc;

  Epi ep;
  ep.out = e->out; ep.ldo = e->ldo; ep.out_is_f32 = e->out_is_f32; ep.act = e->act;
  ep.bias = e->bias; ep.gate = e->gate; ep.residual = e->residual; ep.ldr = e->ldr;
  ep.sumsq = e->sumsq; ep.sumsq_groups = e->sumsq_groups; ep.sumsq_group_cols = e->sumsq_group_cols; ep.sumsq_parts = e->sumsq_parts;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(gemm_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          SMEM_BYTES);
    if (ce != cudaSuccess) {
      set_last_error("svi_gemm_bf16: cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce));
      return SVI_ERR_LAUNCH;
    }
    attr_set = true;
  }
  const int num_tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const int sms = sm_count();
  if (sms <= 0) return SVI_ERR_DRIVER;
  const int grid = num_tiles < sms ? num_tiles : sms;
  gemm_bf16_kernel<<<grid, NUM_THREADS, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(ta, tb, M, N, K, ep);
  SVI_CUDA_LAUNCH_CHECK("svi_gemm_bf16");
  return SVI_OK;
}
