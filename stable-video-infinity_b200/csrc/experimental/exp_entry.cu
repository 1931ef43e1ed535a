// C entry points of the EXPERIMENTAL library (lib/libsvi_b200_exp.so, `make exp`): kernels that are candidates for the
// product but have not earned it yet.  Never loaded by the package; tools/gpu_check.py section `perf_attn4` times them.
#include "../common.cuh"
#include "../../../include/svi_b200.h"

namespace svi {
namespace attn4 {
int launch_attn4(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O, int64_t ldo,
                 int32_t Lq, int32_t Lk, int32_t num_heads, float scale, int32_t accumulate, const uint32_t* kv_flags,
                 uint32_t kv_epoch, int kv_chunk_rows, int kv_self_chunk, void* workspace, size_t workspace_bytes, void* stream,
                 const char* who);
}
}  // namespace svi

extern "C" int svi_exp_attn4(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                             int64_t ldo, int32_t Lq, int32_t Lk, int32_t num_heads, float scale, void* workspace,
                             size_t workspace_bytes, void* stream) {
  return svi::attn4::launch_attn4(Q, ldq, K, ldk, V, ldv, O, ldo, Lq, Lk, num_heads, scale, 0, nullptr, 0, 1, 0, workspace,
                                  workspace_bytes, stream, "svi_exp_attn4");
}
