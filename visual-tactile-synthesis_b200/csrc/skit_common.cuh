// Shared helpers for the skit_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/skit_b200.h"

namespace skit {

// ---- error plumbing (C-ABI functions return an int code; text via skit_last_error) ----
void set_error(const char* fmt, ...);
int  check_launch(const char* what);

#define SKIT_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            skit::set_error(__VA_ARGS__);       \
            return SKIT_ERR_INVALID;            \
        }                                       \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

// Map a coordinate of a padded axis back to the source axis.
// mode: SKIT_PAD_ZERO -> -1 outside; REFLECT (no edge repeat); REPLICATE (clamp).
__host__ __device__ inline int pad_src(int p, int pad, int n, int mode) {
    int s = p - pad;
    if (s >= 0 && s < n) return s;
    if (mode == SKIT_PAD_REFLECT) {
        if (s < 0) s = -s;
        if (s >= n) s = 2 * (n - 1) - s;
        return s;
    }
    if (mode == SKIT_PAD_REPLICATE) return s < 0 ? 0 : n - 1;
    return -1;
}

__device__ inline float act_fwd(float v, int act) {
    if (act == SKIT_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == SKIT_ACT_LRELU) return v > 0.f ? v : 0.2f * v;
    return v;
}
// derivative expressed on the pre-activation value
__device__ inline float act_grad(float pre, int act) {
    if (act == SKIT_ACT_RELU) return pre > 0.f ? 1.f : 0.f;
    if (act == SKIT_ACT_LRELU) return pre > 0.f ? 1.f : 0.2f;
    return 1.f;
}

// fp32 -> (bf16 hi, bf16 lo) with hi + lo == x to ~2^-17 relative.
__device__ inline void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ inline float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ inline double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl() may be scheduled while the kernel in front of it
// in the stream is still draining; everything it does before pdl_wait() (barrier init, TMEM allocation, descriptor prefetch, index
// arithmetic — nothing that touches global memory) overlaps that tail, pdl_wait() returns once the predecessor has completed and its
// writes are visible.  pdl_trigger() in the predecessor lets the dependent grid be scheduled as soon as every CTA of this grid has
// started (without it: when the grid has completed, i.e. no overlap, still correct).  Both are no-ops in a normally launched kernel.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifdef SKIT_PDL_EARLY_TRIGGER
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
// measured: triggering at kernel entry made the 768^2 step slower (27.1 vs 25.9 ms) — the early-scheduled dependents sit in their
// wait holding SM slots the side-stream kernels (weight gradients, discriminator branches) would have used
__device__ __forceinline__ void pdl_trigger() {}
#endif

bool pdl_enabled();      // SKIT_PDL=1 (default off: measured no gain on the train step, see below and DESIGN.md section 10); prep.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace skit
