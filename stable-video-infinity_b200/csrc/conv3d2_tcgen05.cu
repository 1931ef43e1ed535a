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
// Roles per CTA: warp 0 TMA producer, warp 1 MMA issuer (even CTA) + TMEM owner, warps 2-9 epilogue (two per TMEM lane
// quadrant, half of the output channels each; the pixel's sum of squares for the fused RMS norm is exchanged through
// shared memory).  Epilogue semantics are those of conv3d_tcgen05.cu.
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
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
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
// 1..4 K = 16 steps of the M = 256 pair MMA in ONE asm block (whole warp calls with uniform operands; one elected lane issues):
// a C++ loop of single MMAs costs an elect + predicate + descriptor moves per 48-cycle MMA at N = 96
#define SVI_MMA2_HEAD                      \
  "{\n"                                    \
  ".reg .pred p, q, t;\n"                  \
  ".reg .b64 da, db;\n"                    \
  ".reg .b32 a1, b1;\n"                    \
  "setp.ne.b32 p, %6, 0;\n"                \
  "setp.eq.b32 t, 0, 0;\n"                 \
  "elect.sync _|q, 0xffffffff;\n"          \
  "mov.b64 da, {%1, %2};\n"                \
  "mov.b64 db, {%3, %4};\n"                \
  "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
#define SVI_MMA2_STEP(off)                 \
  "add.u32 a1, %1, " #off ";\n"            \
  "add.u32 b1, %3, " #off ";\n"            \
  "mov.b64 da, {a1, %2};\n"                \
  "mov.b64 db, {b1, %4};\n"                \
  "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, t;\n"
#define SVI_MMA2_ARGS \
  ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first) : "memory"
__device__ __forceinline__ void mma2_ss_n(int nk, uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t acc_first) {
  if (nk == 4)
    asm volatile(SVI_MMA2_HEAD SVI_MMA2_STEP(2) SVI_MMA2_STEP(4) SVI_MMA2_STEP(6) "}\n" SVI_MMA2_ARGS);
  else if (nk == 2)
    asm volatile(SVI_MMA2_HEAD SVI_MMA2_STEP(2) "}\n" SVI_MMA2_ARGS);
  else if (nk == 3)
    asm volatile(SVI_MMA2_HEAD SVI_MMA2_STEP(2) SVI_MMA2_STEP(4) "}\n" SVI_MMA2_ARGS);
  else
    asm volatile(SVI_MMA2_HEAD "}\n" SVI_MMA2_ARGS);
}
#undef SVI_MMA2_HEAD
#undef SVI_MMA2_STEP
#undef SVI_MMA2_ARGS
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
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

struct Params {
  svi::conv::Params c;          // geometry + epilogue of the single-CTA kernel (BW / BH / stages unused here)
  int stage_bytes, stages;
  int b_bytes;                  // one weight half-tile: (BN / 2) rows x 128 B
  int nk_last;                  // K = 16 steps of the last channel chunk that hold real channels
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
conv2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
             const __grid_constant__ Params pp) {
  const svi::conv::Params& p = pp.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 1023u) & ~1023u;
  enum : uint32_t { FULL = 0, EMPTY = MAX_STAGES, TMEM_FULL = 2 * MAX_STAGES, TMEM_EMPTY = 2 * MAX_STAGES + 2, NUM_BARS = 2 * MAX_STAGES + 4 };
  auto bar = [&](uint32_t n) { return sbase + RING_BYTES + 8u * n; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_raw + (sbase - sraw) + RING_BYTES + 8 * NUM_BARS);
  float* ssq_x = reinterpret_cast<float*>(smem_raw + (sbase - sraw) + RING_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int STAGES = pp.stages;
  const int tiles_w = (p.W + BM - 1) / BM;
  const int row_pairs = (p.H + 1) / 2;
  const int num_n = (p.C_out + p.BN - 1) / p.BN;
  const int tiles_per_frame = row_pairs * tiles_w;
  const int num_tiles = p.T * tiles_per_frame * num_n;
  const int BNH = p.BN / 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (uint32_t i = 0; i < MAX_STAGES; ++i) {
        mbar_init_a(bar(FULL + i), 1);
        mbar_init_a(bar(EMPTY + i), 1);
      }
      for (uint32_t i = 0; i < 2; ++i) {
        mbar_init_a(bar(TMEM_FULL + i), 1);
        mbar_init_a(bar(TMEM_EMPTY + i), 2 * EPI_WARPS);
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
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tile -> (n block fastest, then 128-pixel segment, then row pair, then frame)
  auto decode_tile = [&](int tile, int& t, int& h0, int& w0, int& n_blk) {
    n_blk = tile % num_n;
    int r = tile / num_n;
    const int sp = r % tiles_per_frame;
    t = r / tiles_per_frame;
    h0 = (sp / tiles_w) * 2;
    w0 = (sp % tiles_w) * BM;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_tx = A_BYTES + p.kw * pp.b_bytes;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        int t, h0, w0, n_blk;
        decode_tile(tile, t, h0, w0, n_blk);
        const int h = h0 + (int)rank;
        const int wrow = n_blk * p.BN + (int)rank * BNH;
        for (int a = 0; a < p.kt; ++a) {
          const int slot = p.slot[t][a];
          for (int b = 0; b < p.kh; ++b) {
            for (int cc = 0; cc < p.cin_chunks; ++cc) {
              mbar_wait_a(bar(EMPTY + stage), phase ^ 1);
              if (rank == 0) mbar_expect_tx_a(bar(FULL + stage), 2 * stage_tx);
              const uint32_t sa = sbase + stage * pp.stage_bytes;
              const uint32_t full_even = bar(FULL + stage) & PEER_MASK;
              tma_load_4d_2sm(sa, &tmap_x, full_even, cc * BK, w0 - p.pad_w, h + b - p.pad_h, slot);
              const int kcol0 = ((a * p.kh + b) * p.kw * p.cin_chunks + cc) * BK;
              for (int c = 0; c < p.kw; ++c)
                tma_load_2d_2sm(sa + A_BYTES + c * pp.b_bytes, &tmap_w, full_even, kcol0 + c * p.cin_chunks * BK, wrow);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * BM, p.BN, 0, 0);
      constexpr uint32_t hi_kmaj = smem_desc_hi(1024, 2);  // SBO 1024 B, 128B swizzle, K-major
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const int groups = p.kt * p.kh;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait_a(bar(TMEM_EMPTY + acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * MAX_BN;
        uint32_t accumulate = 0;
        for (int g = 0; g < groups; ++g) {
          for (int cc = 0; cc < p.cin_chunks; ++cc) {
            mbar_wait_a(bar(FULL + stage), phase);
            tc_fence_after();
            const uint32_t sa = sbase + stage * pp.stage_bytes;
            const int nk = cc == p.cin_chunks - 1 ? pp.nk_last : 4;
            for (int c = 0; c < p.kw; ++c) {
              // tap c of this row = the same box read c pixels (rows of 128 B) further in.  The 128B swizzle is a function of
              // the absolute shared-memory address (TMA wrote the box into a 1024-byte aligned stage), so a start address that is
              // not a multiple of 8 rows needs nothing else: the descriptor's base-offset field stays 0 (measured on B200:
              // with the row phase in that field the results are wrong, profiles/r02_c6_pair_probe.log)
              const uint32_t a_lo = smem_desc_lo(sa + c * 128, 16);
              const uint32_t b_lo = smem_desc_lo(sa + A_BYTES + c * pp.b_bytes, 16);
              mma2_ss_n(nk, d_tmem, a_lo, hi_kmaj, b_lo, hi_kmaj, idesc, accumulate);
              accumulate = 1;
            }
            commit2_multicast(bar(EMPTY + stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        commit2_multicast(bar(TMEM_FULL + acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quad * 32 + lane;
    const int cols_half = p.BN / 2;           // multiple of 16 (launch side)
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int t, h0, w0, n_blk;
      decode_tile(tile, t, h0, w0, n_blk);
      const int h = h0 + (int)rank, w = w0 + row_in_tile;
      const bool ok = (h < p.H) && (w < p.W);
      const long long pix = (long long)h * p.W + w;
      const int col0 = n_blk * p.BN + half * cols_half;
      svi::conv::residual_prefetch(p, cols_half, col0, t, pix, ok);     // while the K loop of this tile still runs
      mbar_wait_a(bar(TMEM_FULL + acc), acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + acc * MAX_BN + half * cols_half + (static_cast<uint32_t>(quad * 32) << 16);
      float ssq = svi::conv::epilogue_pass1(p, t_base, cols_half / 16, col0, t, pix, ok);
      if (p.next_ring) {
        if (p.next_gamma) {
          float* sx = ssq_x + acc * 2 * BM;
          sx[half * BM + row_in_tile] = ssq;
          named_bar_sync(1 + quad, 64);
          ssq = sx[row_in_tile] + sx[BM + row_in_tile];
        }
        svi::conv::epilogue_pass2(p, t_base, cols_half / 16, col0, t, pix, ok, ssq);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar(TMEM_EMPTY + acc), 0);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
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
