// Pieces of the GEMM epilogue shared by gemm_tcgen05.cu (single CTA) and gemm2_tcgen05.cu (CTA pairs).
#pragma once
#include "common.cuh"
#include "../../include/svi_b200.h"

namespace svi {

// Activation of 8 accumulator values with the switch OUTSIDE the element loop.  Written out by hand: with the per-element
// form `v[j] = apply_act(v[j], act)` the compiler's loop unswitching is a heuristic — an unrelated change to the epilogue
// flipped it and every element got its own branch + serial MUFU chain (ffn.0, GELU: 700 -> 1570 us, profiles/r02_c9_bench.err).
__device__ __forceinline__ void apply_act8(float (&v)[8], int act) {
  if (act == SVI_ACT_GELU_TANH) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gelu_tanh(v[j]);
  } else if (act == SVI_ACT_SILU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = silu(v[j]);
  } else if (act == SVI_ACT_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
  } else if (act == SVI_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
  }
}

// Row sums of squares of the epilogue value per column group (the full-width q / k RMS norms, wan_video_dit.py:150-151).
// sumsq_parts == 0: atomicAdd into [M, groups] — the order of the adds varies from run to run, and with it the last bit of
// the sums.  sumsq_parts == P > 0 (P = group_cols / 128): every 128-column segment STORES its partial sum into
// [M, groups * P]; the consumers add a group's P partials in index order, so the result is bit-reproducible and the buffer
// needs no zeroing.  `key` = the segment (or group) the running sum `ss` belongs to.
template <class E>
__device__ __forceinline__ void sumsq_flush(const E& ep, long long row, float ss, int key) {
  if (key < 0) return;
  if (ep.sumsq_parts) {
    const int segs = ep.sumsq_groups * ep.sumsq_parts;
    if (key < segs) ep.sumsq[row * segs + key] = ss;
  } else if (key < ep.sumsq_groups) {
    atomicAdd(&ep.sumsq[row * ep.sumsq_groups + key], ss);
  }
}
template <class E>
__device__ __forceinline__ void sumsq_step(const E& ep, int n0, long long row, float& ss, int& key) {
  const int k = ep.sumsq_parts ? (n0 >> 7) : n0 / ep.sumsq_group_cols;
  if (k != key) {
    sumsq_flush(ep, row, ss, key);
    ss = 0.f;
    key = k;
  }
}

}  // namespace svi
