// Generic fp32 SIMT GEMM with arbitrary operand strides:  C[m,n] = sum_k A(m,k) * B(k,n) (+ bias[n]).
// This is the arithmetic of the fp32 ("exact") mode: true fp32 FFMA accumulation, never TF32, so that
// top-K indices can be compared bit-for-bit with the reference.  It serves the small dense
// layers (fc_mu/fc_logvar, projection MLP, deeper hidden layers; reference src/ml/model.py:90-95,114,126-127)
// forward and backward, and the materialised-score path (decode(), src/ml/model.py:198).
#include "common.cuh"

namespace hvae {

constexpr int GBM = 128, GBN = 64, GBK = 16;

__global__ void __launch_bounds__(256) gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, int64_t a_rs,
                                                       int64_t a_cs, const float* __restrict__ Bm, int64_t b_rs, int64_t b_cs,
                                                       float* __restrict__ C, int64_t ldc, const float* __restrict__ bias,
                                                       float alpha) {
    __shared__ __align__(16) float As[GBK][GBM + 4];
    __shared__ __align__(16) float Bs[GBK][GBN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * GBN;
    const int ty = tid >> 4, tx = tid & 15;
    const bool a_kc = (a_cs == 1), b_kc = (b_rs == 1 && b_cs != 1);

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    float ra[8], rb[4];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int m, k;
            if (a_kc) { k = tid & 15; m = (tid >> 4) + 16 * i; }
            else { m = tid & 127; k = (tid >> 7) + 2 * i; }
            const int gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < M && gk < K) ? A[gm * a_rs + gk * a_cs] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int n, k;
            if (b_kc) { k = tid & 15; n = (tid >> 4) + 16 * i; }
            else { n = tid & 63; k = (tid >> 6) + 4 * i; }
            const int gn = n0 + n, gk = k0 + k;
            rb[i] = (gn < N && gk < K) ? Bm[gk * b_rs + gn * b_cs] : 0.f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int m, k;
            if (a_kc) { k = tid & 15; m = (tid >> 4) + 16 * i; }
            else { m = tid & 127; k = (tid >> 7) + 2 * i; }
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int n, k;
            if (b_kc) { k = tid & 15; n = (tid >> 4) + 16 * i; }
            else { n = tid & 63; k = (tid >> 6) + 4 * i; }
            Bs[k][n] = rb[i];
        }
    };

    load_tiles(0);
    for (int k0 = 0; k0 < K; k0 += GBK) {
        store_tiles();
        __syncthreads();
        if (k0 + GBK < K) load_tiles(k0 + GBK);
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int gm = m0 + ty * 8 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j] * alpha;
            if (bias) v += bias[gn];
            C[gm * ldc + gn] = v;
        }
    }
}

}  // namespace hvae

extern "C" int hvae_gemm_f32(int M, int N, int K, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs,
                             int64_t b_cs, float* C, int64_t ldc, const float* bias, float alpha, void* stream) {
    if (M == 0 || N == 0) return 0;
    dim3 grid(hvae::ceil_div(N, hvae::GBN), hvae::ceil_div(M, hvae::GBM));
    HVAE_REQUIRE(grid.y <= 65535, "gemm_f32: M=%d too large for one launch", M);
    hvae::gemm_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, alpha);
    HVAE_LAUNCH_CHECK("gemm_f32");
    return 0;
}
