// Microbenchmark: raw tcgen05.mma issue/execute rate per instruction shape on one CTA per SM (no TMA, no epilogue).
// Prints cycles per MMA and the implied fraction of the 8192 FLOP/clk/SM bf16 peak.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "../../stable-video-infinity_b200/csrc/common.cuh"

namespace svi { void set_last_error(const char*, ...) {} }
using namespace svi;

// mode 0: SS (A, B from smem, K-major both); mode 1: TS (A from TMEM, B MN-major from smem)
template <int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int N, int iters, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_ptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&tmem_ptr, 512); tmem_relinquish(); }
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  // fill smem with something (values irrelevant)
  for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sbase + i * 4), "r"(0x3c003c00u));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, MODE == 1 ? 1 : 0);
    constexpr uint32_t hi = smem_desc_hi(1024, 2);
    const uint32_t a_lo = smem_desc_lo(sbase, 16);
    const uint32_t b_lo = smem_desc_lo(sbase + 32768, MODE == 1 ? 16384 : 16);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // 8 K-steps per iteration, alternating two accumulators like a real kernel
      const uint32_t d = (it & 1) * 256;
      if (MODE == 0) {
        tc_mma_ss_k4(d, a_lo, hi, b_lo, hi, idesc, 0);
        tc_mma_ss_k4(d, a_lo + 1024, hi, b_lo + 2048, hi, idesc, 1);
      } else {
        tc_mma_ts_k4(d, 448, b_lo, hi, idesc, 0);
        tc_mma_ts_k4(d, 448 + 32, b_lo + 512, hi, idesc, 1);
      }
    }
    tc_commit_p(0, &bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) *out_cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(0, 512); }
}

int main() {
  long long* d_c;
  cudaMalloc(&d_c, 8);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int N : {64, 128, 256}) {
      for (int grid : {1, 148}) {
        for (int rep = 0; rep < 2; ++rep) {
          if (mode == 0) mma_rate_kernel<0><<<grid, 128, smem>>>(N, iters, d_c);
          else mma_rate_kernel<1><<<grid, 128, smem>>>(N, iters, d_c);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long c;
        cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
        double per = (double)c / (iters * 8.0);
        double flop_per_clk = 2.0 * 128 * N * 16 / per;
        printf("%s M=128 N=%3d K=16 grid=%3d: %.1f cycles/MMA  -> %.0f FLOP/clk/SM = %.1f%% of 8192\n",
               mode == 0 ? "SS" : "TS", N, grid, per, flop_per_clk, 100.0 * flop_per_clk / 8192.0);
      }
    }
  return 0;
}
