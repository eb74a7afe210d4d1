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
