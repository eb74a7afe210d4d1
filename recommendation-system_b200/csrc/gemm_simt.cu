// Generic fp32 SIMT GEMM with arbitrary operand strides:  C[m,n] = alpha * sum_k A(m,k) * B(k,n) (+ bias[n]).
// This is the arithmetic of the fp32 ("exact") mode: true fp32 FFMA accumulation, never TF32, so that
// top-K indices can be compared bit-for-bit with the reference.  It serves the small dense
// layers (fc_mu/fc_logvar, projection MLP, deeper hidden layers; reference src/ml/model.py:90-95,114,126-127)
// forward and backward, and the materialised-score path (decode(), src/ml/model.py:198).
// Two tile shapes: 128x64 for large problems, 32x64 when the large tile would leave most of the 148 SMs idle
// (the MLP GEMMs of a 512-user batch are 400-600 wide: latency-bound, so CTA count matters more than reuse; the
// small tile stages 64-deep k-blocks so that one prefetch covers the HBM latency).
#include "common.cuh"

namespace hvae {

template <int BM, int BN, int TM, int TN, int GBK>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, int64_t a_rs,
                                                                         int64_t a_cs, const float* __restrict__ Bm, int64_t b_rs,
                                                                         int64_t b_cs, float* __restrict__ C, int64_t ldc,
                                                                         const float* __restrict__ bias, float alpha) {
    pdl_prologue();
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int LA = BM * GBK / NT, LB = BN * GBK / NT;     // elements each thread stages per k-block
    static_assert(LA >= 1 && LB >= 1 && TM % 4 == 0 && TN == 4, "tile configuration");
    __shared__ __align__(16) float As[GBK][BM + 4];
    __shared__ __align__(16) float Bs[GBK][BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int ty = tid / (BN / TN), tx = tid % (BN / TN);
    const bool a_kc = (a_cs == 1), b_kc = (b_rs == 1 && b_cs != 1);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float ra[LA], rb[LB];
    auto a_pos = [&](int i, int& m, int& k) {
        if (a_kc) { k = tid % GBK; m = tid / GBK + (NT / GBK) * i; }
        else { m = tid % BM; k = tid / BM + (NT / BM) * i; }
    };
    auto b_pos = [&](int i, int& n, int& k) {
        if (b_kc) { k = tid % GBK; n = tid / GBK + (NT / GBK) * i; }
        else { n = tid % BN; k = tid / BN + (NT / BN) * i; }
    };
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            int m, k;
            a_pos(i, m, k);
            const int gm = m0 + m, gk = k0 + k;
            ra[i] = (gm < M && gk < K) ? A[gm * a_rs + gk * a_cs] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            int n, k;
            b_pos(i, n, k);
            const int gn = n0 + n, gk = k0 + k;
            rb[i] = (gn < N && gk < K) ? Bm[gk * b_rs + gn * b_cs] : 0.f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < LA; ++i) { int m, k; a_pos(i, m, k); As[k][m] = ra[i]; }
#pragma unroll
        for (int i = 0; i < LB; ++i) { int n, k; b_pos(i, n, k); Bs[k][n] = rb[i]; }
    };

    load_tiles(0);
    for (int k0 = 0; k0 < K; k0 += GBK) {
        store_tiles();
        __syncthreads();
        if (k0 + GBK < K) load_tiles(k0 + GBK);
#pragma unroll
        for (int k = 0; k < GBK; ++k) {
            float av[TM];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * TM + i]);
                av[i] = a.x; av[i + 1] = a.y; av[i + 2] = a.z; av[i + 3] = a.w;
            }
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty * TM + i;
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
    using namespace hvae;
    if (M == 0 || N == 0) return 0;
    const int big_ctas = ceil_div(N, 64) * ceil_div(M, 128);
    if (big_ctas >= 2 * kNumSMs) {
        dim3 grid(ceil_div(N, 64), ceil_div(M, 128));
        HVAE_REQUIRE(grid.y <= 65535, "gemm_f32: M=%d too large for one launch", M);
        launch_pdl(gemm_f32_kernel<128, 64, 8, 4, 16>, grid, 256, 0, (cudaStream_t)stream, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, alpha);
    } else {
        dim3 grid(ceil_div(N, 64), ceil_div(M, 32));
        HVAE_REQUIRE(grid.y <= 65535, "gemm_f32: M=%d too large for one launch", M);
        launch_pdl(gemm_f32_kernel<32, 64, 4, 4, 64>, grid, 128, 0, (cudaStream_t)stream, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, bias, alpha);
    }
    HVAE_LAUNCH_CHECK("gemm_f32");
    return 0;
}
