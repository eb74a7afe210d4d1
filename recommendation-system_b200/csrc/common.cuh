// Shared device helpers for the hvae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Records a message retrievable through hvae_last_error() and returns a non-zero status.
int hvae_fail(const char* fmt, ...);

#define HVAE_LAUNCH_CHECK(what)                                                       \
    do {                                                                              \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) return hvae_fail("%s: %s", what, cudaGetErrorString(_e)); \
    } while (0)

#define HVAE_REQUIRE(cond, ...)                    \
    do {                                           \
        if (!(cond)) return hvae_fail(__VA_ARGS__); \
    } while (0)

#define HVAE_CUDA(call)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) return hvae_fail("%s: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

namespace hvae {

constexpr int kNumSMs = 148;

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// Every kernel of the library is launched with the programmatic-stream-serialization attribute and starts with
// pdl_prologue(): "my dependents may start launching" + "wait until the kernel(s) before me have completed and flushed".
// Inside a captured step this overlaps the launch latency / prologue of kernel N+1 with the execution of kernel N (the
// tensor-core kernels do their barrier init, TMEM allocation and descriptor prefetch before the wait).  Without the
// attribute both instructions are no-ops.  HVAE_NO_PDL=1 disables the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }

bool pdl_enabled();   // capi.cu

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact-erf GELU and its derivative (nn.GELU() default, reference src/ml/model.py:92,116)
__device__ __forceinline__ float gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace hvae
