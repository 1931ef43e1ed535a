// Shared by the two convolution kernels (conv3d_tcgen05.cu: single CTA, one box per tap; conv3d2_tcgen05.cu: CTA pairs,
// input window reused across the horizontal taps): launch parameters and the accumulator epilogue.
#pragma once
#include "common.cuh"
#include "../../include/svi_b200.h"

namespace svi {
namespace conv {

struct Params {
  int H, W, T;                 // output height / width / frames of this launch
  int kt, kh, kw;              // kernel extent
  int pad_h, pad_w;            // input coordinate = output coordinate + k - pad
  int slot[4][3];              // ring slot of the input frame for (output frame, k_t)
  int cin_chunks;              // ceil(C_in / 64)
  int C_out, BN, BW, BH;       // BH * BW == 128 (single-CTA kernel)
  float* out;                  // fp32 channels-last [frame][H][W][out_ld]
  long long out_frame_stride;  // elements between output frames
  int out_ld;                  // channels per output pixel in memory
  int n_split;                 // columns >= n_split are written to out + split_offset (column - n_split); 0 = off
  long long split_offset;
  const float* bias;           // [C_out]
  const float* residual;       // same layout as out (frame stride res_frame_stride, ld res_ld) or null
  long long res_frame_stride;
  int res_ld;
  int stages, stage_bytes;     // operand ring of the single-CTA kernel: stage_bytes = A tile + BN x 64 bf16 rounded to 1 KB
  // fused producer of the NEXT conv's input (RMS_norm + SiLU + bf16 staging, wan_video_vae.py:55-70,206-210):
  __nv_bfloat16* next_ring;    // bf16 [slots][H][W][next_ld] or null
  long long next_frame_stride; // elements between ring slots
  int next_ld;
  int next_slot[4];            // ring slot of output frame t
  const float* next_gamma;     // [C_out] RMS-norm weight, null: plain cast
  int next_silu;
  int write_f32;               // 0: the fp32 output is not needed (only the normalised bf16 is consumed)
};

void* encode_tiled_fn();       // cuTensorMapEncodeTiled through the runtime's driver entry point query (null + last error on failure)

// ---- accumulator epilogue ------------------------------------------------------------------------------------------
// One accumulator row = one output pixel (TMEM lane = pixel); a warp walks `n16` groups of 16 columns starting at TMEM
// address t_base = global output channel col0.  The fp32 residual of a pixel is a strided 16-byte read per lane: a load whose
// latency is exposed once per dependent use costs ~1-2 k cycles, and with plain (possibly aliasing) out / residual pointers the
// compiler keeps every load behind the previous store — the row-owner epilogue measured 40 % slower than the K loop it should
// hide behind (profiles/r02_c6_perf_conv.log, 'end' against 'mid').  So: the residual lines are prefetched into L2 before the
// accumulator is waited for (residual_prefetch), and every pair of column groups issues its eight residual loads up front.

__device__ __forceinline__ void residual_prefetch(const Params& p, int n_cols, int col0, int t, long long pix, bool ok) {
  if (!p.residual || !ok) return;
  const char* rp = reinterpret_cast<const char*>(p.residual + (long long)t * p.res_frame_stride + pix * p.res_ld + col0);
  const int bytes = min(n_cols, p.C_out - col0) * 4;
  for (int b = 0; b < bytes; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + b));
}

// loads the residual of column groups c and c + 1 (8 x float4; zero where there is none)
__device__ __forceinline__ void residual_load2(const Params& p, const float* rrow, int n0, bool ok, float4 (&q)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int n = n0 + j * 4;
    q[j] = (rrow && ok && n < p.C_out) ? *reinterpret_cast<const float4*>(rrow + n) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Pass 1:  v = acc + bias (+ residual) -> fp32 out (if write_f32);  returns the sum of v^2 over the columns seen (the fused RMS
// norm of the next conv's input needs it over the whole channel vector).
__device__ __forceinline__ float epilogue_pass1(const Params& p, uint32_t t_base, int n16, int col0, int t, long long pix, bool ok) {
  float ssq = 0.f;
  const float* rrow = p.residual ? p.residual + (long long)t * p.res_frame_stride + pix * p.res_ld : nullptr;
#pragma unroll 1
  for (int c = 0; c < n16; c += 2) {
    const int n0 = col0 + c * 16;
    if (n0 >= p.C_out) break;
    float4 q[8];
    residual_load2(p, rrow, n0, ok, q);
    const bool two = c + 1 < n16 && n0 + 16 < p.C_out;
    uint32_t r[32];
    tmem_ld16(t_base + c * 16, *reinterpret_cast<uint32_t(*)[16]>(r));
    if (two) tmem_ld16(t_base + c * 16 + 16, *reinterpret_cast<uint32_t(*)[16]>(r + 16));
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const int n = n0 + j4 * 4;
        if (n >= p.C_out || (j4 >= 4 && !two)) break;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r[j4 * 4 + j]);
        if (p.bias) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
        v[0] += q[j4].x; v[1] += q[j4].y; v[2] += q[j4].z; v[3] += q[j4].w;
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
  return ssq;
}

// Pass 2 over the same accumulator columns (still in TMEM): y = silu(v / max(||v||, 1e-12) * sqrt(C) * gamma) -> bf16 into the
// next conv's input ring; ssq = sum of v^2 over the pixel's WHOLE channel vector.
__device__ __forceinline__ void epilogue_pass2(const Params& p, uint32_t t_base, int n16, int col0, int t, long long pix, bool ok,
                                               float ssq) {
  const float mul = p.next_gamma ? sqrtf((float)p.C_out) / fmaxf(sqrtf(ssq), 1e-12f) : 1.f;
  __nv_bfloat16* nrow = p.next_ring + (long long)p.next_slot[t] * p.next_frame_stride + pix * p.next_ld;
  const float* rrow = p.residual ? p.residual + (long long)t * p.res_frame_stride + pix * p.res_ld : nullptr;
#pragma unroll 1
  for (int c = 0; c < n16; c += 2) {
    const int n0 = col0 + c * 16;
    if (n0 >= p.C_out) break;
    float4 q[8];
    residual_load2(p, rrow, n0, ok, q);
    const bool two = c + 1 < n16 && n0 + 16 < p.C_out;
    uint32_t r[32];
    tmem_ld16(t_base + c * 16, *reinterpret_cast<uint32_t(*)[16]>(r));
    if (two) tmem_ld16(t_base + c * 16 + 16, *reinterpret_cast<uint32_t(*)[16]>(r + 16));
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const int n = n0 + j4 * 4;
        if (n >= p.C_out || (j4 >= 4 && !two)) break;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r[j4 * 4 + j]);
        if (p.bias) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n));
          v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
        }
        v[0] += q[j4].x; v[1] += q[j4].y; v[2] += q[j4].z; v[3] += q[j4].w;
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

}  // namespace conv

namespace conv2 {
constexpr int AUTO_MIN_W = 192;   // below this the 128-pixel row tiles of the pair kernel waste more than the reuse gains (measured: profiles/)
int pair_bn(const svi_conv_desc* d, int BN);   // N tile of the pair kernel, 0: not applicable
int launch(const svi_conv_desc* d, const svi::conv::Params& base, int BN, cudaStream_t stream);
}  // namespace conv2
}  // namespace svi
