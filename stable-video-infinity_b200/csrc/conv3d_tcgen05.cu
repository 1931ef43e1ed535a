// Causal 3-D / 2-D convolution of the Wan VAE as an implicit GEMM on tcgen05 (sm_100a).
//
// Activations are channels-last bf16 "frame rings"  [slots][H][W][C];  one CTA tile = 128 output pixels
// (a BH x BW patch of one frame, TMEM lane = pixel) x BN output channels.  The K loop walks
// (k_t, k_h, k_w, 64-channel chunk): for every step ONE 4-D TMA box {64ch, BW, BH, 1 frame} of the input ring,
// shifted by the tap offset, lands in shared memory as a canonical 128-row x 128-byte swizzled K-major
// operand tile — spatial zero padding comes for free from TMA out-of-bounds zero fill, temporal causality
// from the (output frame, k_t) -> ring slot table (history frames stay in the ring: no F.pad / torch.cat
// cache copies as in the reference, wan_video_vae.py:44-52, 214-232).  Weights are pre-packed
// [C_out, (k_t,k_h,k_w,C_in padded to 64)] and stream through a 2-D TMA like a GEMM B operand.
// Warp roles / pipelines are those of gemm_tcgen05.cu (operand ring of 4..7 stages depending on the N tile).  Epilogue:
// + bias, + fp32 residual (ResidualBlock skip), optional channel split into two frames (upsample3d time_conv,
// wan_video_vae.py:153-156), fp32 NHWC out — and, fused, the NEXT conv's input: RMS_norm + SiLU (:55-70, 206-210) of the
// output pixel (its whole channel vector sits in one accumulator row) written as bf16 straight into the next conv's frame
// ring, so the activation between the two convs of a ResidualBlock never exists in fp32 and no staging pass runs.
#include "common.cuh"
#include "../../include/svi_b200.h"
#include "conv3d_common.cuh"

namespace svi {
namespace conv {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 8;
constexpr int MAX_BN = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int RING_BYTES = 4 * (A_STAGE_BYTES + MAX_BN * BK * 2);   // 192 KB of operand stages: 4 at BN = 256 ... 7 at BN = 96
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = RING_BYTES + 1024 + 256;


__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * MAX_STAGES;
  uint64_t* tmem_empty_bar = bars + 2 * MAX_STAGES + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  const int STAGES = p.stages;
  const int STAGE_BYTES = p.stage_bytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_w = (p.W + p.BW - 1) / p.BW;
  const int tiles_h = (p.H + p.BH - 1) / p.BH;
  const int num_n = (p.C_out + p.BN - 1) / p.BN;
  const int tiles_per_frame = tiles_h * tiles_w;
  const int num_tiles = p.T * tiles_per_frame * num_n;
  const int num_k = p.kt * p.kh * p.kw * p.cin_chunks;
  const uint32_t stage_tx = A_STAGE_BYTES + p.BN * BK * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < MAX_STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full_bar[i], 1);
        mbar_init(&tmem_empty_bar[i], 4);
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

  // tile index -> (n block fastest, then spatial tile, then frame): CTAs running together share the input patch
  auto decode_tile = [&](int tile, int& t, int& h0, int& w0, int& n_blk) {
    n_blk = tile % num_n;
    int r = tile / num_n;
    const int sp = r % tiles_per_frame;
    t = r / tiles_per_frame;
    h0 = (sp / tiles_w) * p.BH;
    w0 = (sp % tiles_w) * p.BW;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int t, h0, w0, n_blk;
        decode_tile(tile, t, h0, w0, n_blk);
        int kcol = 0;
        for (int a = 0; a < p.kt; ++a) {
          const int slot = p.slot[t][a];
          for (int b = 0; b < p.kh; ++b) {
            for (int c = 0; c < p.kw; ++c) {
              for (int cc = 0; cc < p.cin_chunks; ++cc) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                mbar_expect_tx(&full_bar[stage], stage_tx);
                uint8_t* sa = smem + stage * STAGE_BYTES;
                tma_load_4d(sa, &tmap_x, &full_bar[stage], cc * BK, w0 + c - p.pad_w, h0 + b - p.pad_h, slot);
                tma_load_2d(sa + A_STAGE_BYTES, &tmap_w, &full_bar[stage], kcol, n_blk * p.BN);
                kcol += BK;
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(BM, p.BN, 0, 0);
    constexpr uint32_t hi_kmaj = smem_desc_hi(1024, 2);  // SBO 1024 B, 128B swizzle, K-major
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * MAX_BN;
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
    const int quad = warp & 3;
    const int row_in_tile = quad * 32 + lane;
    const int hl = row_in_tile / p.BW, wl = row_in_tile % p.BW;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int t, h0, w0, n_blk;
      decode_tile(tile, t, h0, w0, n_blk);
      const int h = h0 + hl, w = w0 + wl;
      const bool ok = (h < p.H) && (w < p.W);
      const long long pix = (long long)h * p.W + w;
      residual_prefetch(p, p.BN, n_blk * p.BN, t, pix, ok);     // while the K loop of this tile still runs
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + acc * MAX_BN + (static_cast<uint32_t>(quad * 32) << 16);
      // pass 1: + bias (+ residual) -> fp32 out, sum of squares of the pixel's channels; pass 2 (fused producer of the next
      // conv's input; the launch side guarantees one N tile): RMS norm + SiLU -> bf16 ring   (conv3d_common.cuh)
      const float ssq = epilogue_pass1(p, t_base, p.BN / 16, n_blk * p.BN, t, pix, ok);
      if (p.next_ring) epilogue_pass2(p, t_base, p.BN / 16, 0, t, pix, ok, ssq);
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

void* encode_tiled_fn() {
  static void* fn = nullptr;
  if (!fn) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) {
      set_last_error("svi_conv3d_causal: cuTensorMapEncodeTiled unavailable");
      return nullptr;
    }
    fn = fp;
  }
  return fn;
}

}  // namespace conv
}  // namespace svi

extern "C" int svi_conv3d_causal(const svi_conv_desc* d, void* stream) {
  using namespace svi;
  using namespace svi::conv;
  SVI_REQUIRE(d && d->x_ring && d->w_packed && (d->out || (d->next_ring && !d->write_f32)), "svi_conv3d_causal: null pointer");
  SVI_REQUIRE(d->H > 0 && d->W > 0 && d->T > 0 && d->T <= 4, "svi_conv3d_causal: need 1 <= T <= 4, H, W > 0");
  SVI_REQUIRE(d->kt >= 1 && d->kt <= 3 && d->kh >= 1 && d->kh <= 3 && d->kw >= 1 && d->kw <= 3,
              "svi_conv3d_causal: kernel extents must be in [1,3]");
  SVI_REQUIRE(d->C_in > 0 && d->C_in % 8 == 0, "svi_conv3d_causal: C_in must be a multiple of 8 (got %d)", d->C_in);
  SVI_REQUIRE(d->C_out > 0 && d->C_out % 4 == 0 && d->out_ld % 4 == 0, "svi_conv3d_causal: C_out, out_ld %% 4");
  SVI_REQUIRE(d->ring_slots > 0 && d->in_H > 0 && d->in_W > 0, "svi_conv3d_causal: bad ring geometry");
  SVI_REQUIRE(d->tile_w == 8 || d->tile_w == 16 || d->tile_w == 32 || d->tile_w == 64 || d->tile_w == 128,
              "svi_conv3d_causal: tile_w must be 8, 16, 32, 64 or 128");
  for (int t = 0; t < d->T; ++t)
    for (int a = 0; a < d->kt; ++a)
      SVI_REQUIRE(d->slot[t * 3 + a] >= 0 && d->slot[t * 3 + a] < d->ring_slots, "svi_conv3d_causal: slot out of range");
  const int cin_chunks = (d->C_in + BK - 1) / BK;
  const int cpad = cin_chunks * BK;
  const int ktot = d->kt * d->kh * d->kw * cpad;
  SVI_REQUIRE(d->w_ld >= ktot && d->w_ld % 8 == 0, "svi_conv3d_causal: packed weight ld must be >= kt*kh*kw*ceil64(C_in)");
  int BN = d->C_out <= MAX_BN ? ((d->C_out + 15) / 16) * 16 : 0;
  if (BN == 0) {  // split wide outputs into equal tiles that are multiples of 16
    for (int parts = 2; parts <= 8 && BN == 0; ++parts)
      if (d->C_out % parts == 0 && (d->C_out / parts) % 16 == 0 && d->C_out / parts <= MAX_BN) BN = d->C_out / parts;
    if (BN == 0) BN = MAX_BN;
  }
  SVI_REQUIRE(d->w_rows >= d->C_out, "svi_conv3d_causal: packed weight has fewer rows than C_out");

  Params p;
  p.H = d->H; p.W = d->W; p.T = d->T;
  p.kt = d->kt; p.kh = d->kh; p.kw = d->kw;
  p.pad_h = d->pad_h; p.pad_w = d->pad_w;
  for (int t = 0; t < 4; ++t)
    for (int a = 0; a < 3; ++a) p.slot[t][a] = d->slot[t * 3 + a];
  p.cin_chunks = cin_chunks;
  p.C_out = d->C_out; p.BN = BN; p.BW = d->tile_w; p.BH = BM / d->tile_w;
  p.out = d->out; p.out_frame_stride = d->out_frame_stride; p.out_ld = d->out_ld;
  p.n_split = d->n_split; p.split_offset = d->split_offset;
  p.bias = d->bias; p.residual = d->residual; p.res_frame_stride = d->res_frame_stride; p.res_ld = d->res_ld;
  p.stage_bytes = (A_STAGE_BYTES + BN * BK * 2 + 1023) / 1024 * 1024;
  p.stages = RING_BYTES / p.stage_bytes;
  if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
  p.next_ring = reinterpret_cast<__nv_bfloat16*>(d->next_ring);
  p.next_frame_stride = d->next_frame_stride;
  p.next_ld = d->next_ld;
  for (int t = 0; t < 4; ++t) p.next_slot[t] = d->next_slot[t];
  p.next_gamma = d->next_gamma;
  p.next_silu = d->next_silu;
  p.write_f32 = d->next_ring ? d->write_f32 : 1;
  if (d->next_ring) {
    SVI_REQUIRE(d->C_out <= BN && d->n_split == 0, "svi_conv3d_causal: the fused next-input producer needs one N tile (C_out <= 256) and no channel split");
    SVI_REQUIRE(d->next_ld >= d->C_out && d->next_ld % 4 == 0, "svi_conv3d_causal: next_ld must be >= C_out and a multiple of 4");
    for (int t = 0; t < d->T; ++t) SVI_REQUIRE(d->next_slot[t] >= 0, "svi_conv3d_causal: next_slot must be >= 0");
  }


  // kernel choice: 1 = single CTA, one input box per tap; 2 = CTA pairs with the input window reused across the horizontal
  // taps (conv3d2_tcgen05.cu); 0 = pair kernel where it applies and the image rows are long enough to fill its 128-pixel row tiles
  SVI_REQUIRE(d->variant >= 0 && d->variant <= 2, "svi_conv3d_causal: variant must be 0 (auto), 1 or 2");
  const int pair_bn = svi::conv2::pair_bn(d, BN);
  SVI_REQUIRE(d->variant != 2 || pair_bn > 0, "svi_conv3d_causal: variant 2 needs k_w = 3 and pad_w = 1");
  if (d->variant == 2 || (d->variant == 0 && pair_bn > 0 && d->W >= svi::conv2::AUTO_MIN_W))
    return svi::conv2::launch(d, p, pair_bn, static_cast<cudaStream_t>(stream));

  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(encode_tiled_fn());
  if (!enc) return SVI_ERR_DRIVER;
  const int BW = d->tile_w, BH = BM / BW;
  CUtensorMap tx, tw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->C_in, (cuuint64_t)d->in_W, (cuuint64_t)d->in_H, (cuuint64_t)d->ring_slots};
    cuuint64_t strides[3] = {(cuuint64_t)d->C_in * 2, (cuuint64_t)d->in_W * d->C_in * 2,
                             (cuuint64_t)d->in_H * d->in_W * d->C_in * 2};
    cuuint32_t box[4] = {BK, (cuuint32_t)BW, (cuuint32_t)BH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->x_ring), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("svi_conv3d_causal: input tensor map failed (CUresult %d) C=%d W=%d H=%d slots=%d", (int)r, d->C_in,
                     d->in_W, d->in_H, d->ring_slots);
      return SVI_ERR_DRIVER;
    }
  }
  int rc = make_tmap_2d(&tw, d->w_packed, 2, (uint64_t)ktot, (uint64_t)d->w_rows, (uint64_t)d->w_ld * 2, BK, BN);
  if (rc) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t ce = cudaFuncSetAttribute(conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (ce != cudaSuccess) {
      set_last_error("svi_conv3d_causal: cudaFuncSetAttribute failed: %s", cudaGetErrorString(ce));
      return SVI_ERR_LAUNCH;
    }
    attr_set = true;
  }
  const int tiles = d->T * ((d->H + BH - 1) / BH) * ((d->W + BW - 1) / BW) * ((d->C_out + BN - 1) / BN);
  const int sms = sm_count();
  if (sms <= 0) return SVI_ERR_DRIVER;
  conv_kernel<<<tiles < sms ? tiles : sms, NUM_THREADS, SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(tx, tw, p);
  SVI_CUDA_LAUNCH_CHECK("svi_conv3d_causal");
  return SVI_OK;
}
