// Kernels that only the Mult-VAE baseline needs (reference src/ml/baseline.py:126-231); everything else of that model runs on
// the HybridVAE kernels: the CSR gather-sum for the first Linear, the MLP GEMMs, reparameterise/KL, the materialised fp32
// scoring path, the item-major gradient reduction and the fused Adam.
//   * hvae_mv_input_values : the model's input transform on the SPARSE row -- F.normalize(x, p=2, dim=1) followed by input
//     dropout (baseline.py:151) -- as per-entry values val_j = x_j / max(||x_b||, 1e-12) * keep_j / (1 - p);
//   * hvae_tanh_drop_fwd/bwd: Tanh (+ Dropout) of the hidden layers (baseline.py:137-147); the forward can also append the
//     constant-one column that turns the output layer's bias into one more weight column;
//   * hvae_rows_axpy       : dst[item[s], :] += alpha * src[s, :] for the touched items (the "- x" part of the output layer's
//     weight gradient, which only touches the rows of the batch's items).
#include "common.cuh"
#include "hvae_b200.h"

namespace hvae {

// one warp per batch row: ||x_b||_2 over the row's entries, then the scaled (and masked) values at the entries' global positions
__global__ void __launch_bounds__(256) mv_input_values_kernel(const int64_t* __restrict__ indptr, const float* __restrict__ values,
                                                              const int32_t* __restrict__ rows, int B,
                                                              const uint8_t* __restrict__ keep, const int32_t* __restrict__ keep_ptr,
                                                              float keep_scale, float* __restrict__ out) {
    pdl_prologue();
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const int u = rows ? rows[b] : b;
    const int64_t s = indptr[u], e = indptr[u + 1];
    float q = 0.f;
    for (int64_t j = s + lane; j < e; j += 32) { const float v = values ? values[j] : 1.0f; q = fmaf(v, v, q); }
    q = warp_sum(q);
    const float inv = 1.0f / fmaxf(sqrtf(q), 1e-12f);                 // F.normalize: x / max(||x||, eps), eps = 1e-12
    const int64_t kbase = keep ? keep_ptr[b] : 0;                     // keep flags are stored per batch entry, batch order
    for (int64_t j = s + lane; j < e; j += 32) {
        float v = (values ? values[j] : 1.0f) * inv;
        if (keep) v = keep[kbase + (j - s)] ? v * keep_scale : 0.f;
        out[j] = v;
    }
}

// t = dropout(tanh(q)); pad columns zero; ones_col >= 0: t[:, ones_col] = 1 (bias column of the augmented output layer)
__global__ void tanh_drop_fwd_kernel(const float* __restrict__ q, const uint8_t* __restrict__ mask, float keep_scale, int B, int d,
                                     int ldq, float* __restrict__ t, int ldt, int ones_col) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * ldt) return;
    const int row = i / ldt, j = i - row * ldt;
    float g = 0.f;
    if (j < d) {
        g = tanhf(q[(size_t)row * ldq + j]);
        if (mask) g = mask[(size_t)row * d + j] ? g * keep_scale : 0.f;
    } else if (j == ones_col) {
        g = 1.0f;
    }
    t[i] = g;
}

// dq = dt * mask * keep_scale * (1 - tanh(q)^2)   (dt may alias dq when the leading dimensions agree)
__global__ void tanh_drop_bwd_kernel(const float* dt, int lddt, const float* __restrict__ q, const uint8_t* __restrict__ mask,
                                     float keep_scale, int B, int d, int ldq, float* dq) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * ldq) return;
    const int row = i / ldq, j = i - row * ldq;
    float g = 0.f;
    if (j < d) {
        g = dt[(size_t)row * lddt + j];
        if (mask) g = mask[(size_t)row * d + j] ? g * keep_scale : 0.f;
        const float th = tanhf(q[i]);
        g *= 1.0f - th * th;
    }
    dq[i] = g;
}

__global__ void rows_axpy_kernel(const int32_t* __restrict__ item, const int32_t* __restrict__ n_rows, int max_rows,
                                 const float4* __restrict__ src, int ld4_src, float alpha, float4* __restrict__ dst, int ld4_dst,
                                 int cols4) {
    pdl_prologue();
    const int s = blockIdx.x;
    if (s >= max_rows || s >= *n_rows) return;
    float4* d = dst + (size_t)item[s] * ld4_dst;
    const float4* a = src + (size_t)s * ld4_src;
    for (int c = threadIdx.x; c < cols4; c += blockDim.x) {
        float4 x = d[c];
        const float4 y = a[c];
        x.x = fmaf(alpha, y.x, x.x); x.y = fmaf(alpha, y.y, x.y); x.z = fmaf(alpha, y.z, x.z); x.w = fmaf(alpha, y.w, x.w);
        d[c] = x;
    }
}

}  // namespace hvae

using namespace hvae;

extern "C" {

int hvae_mv_input_values(const int64_t* indptr, const float* values, const int32_t* rows, int B, const uint8_t* keep,
                         const int32_t* keep_ptr, float keep_scale, float* out, void* stream) {
    if (B == 0) return 0;
    HVAE_REQUIRE(!keep || keep_ptr, "mv_input_values: keep flags need keep_ptr (batch-order offsets)");
    launch_pdl(mv_input_values_kernel, ceil_div(B, 8), 256, 0, (cudaStream_t)stream, indptr, values, rows, B, keep, keep_ptr, keep_scale, out);
    HVAE_LAUNCH_CHECK("mv_input_values");
    return 0;
}

int hvae_tanh_drop_fwd(const float* q, const uint8_t* mask, float keep_scale, int B, int d, int ldq, float* t, int ldt, int ones_col,
                       void* stream) {
    if (B == 0) return 0;
    HVAE_REQUIRE(ldt >= d && ldq >= d && ones_col < ldt && (ones_col < 0 || ones_col >= d), "tanh_drop_fwd: bad leading dimensions");
    launch_pdl(tanh_drop_fwd_kernel, ceil_div(B * ldt, 256), 256, 0, (cudaStream_t)stream, q, mask, keep_scale, B, d, ldq, t, ldt, ones_col);
    HVAE_LAUNCH_CHECK("tanh_drop_fwd");
    return 0;
}

int hvae_tanh_drop_bwd(const float* dt, int lddt, const float* q, const uint8_t* mask, float keep_scale, int B, int d, int ldq,
                       float* dq, void* stream) {
    if (B == 0) return 0;
    launch_pdl(tanh_drop_bwd_kernel, ceil_div(B * ldq, 256), 256, 0, (cudaStream_t)stream, dt, lddt, q, mask, keep_scale, B, d, ldq, dq);
    HVAE_LAUNCH_CHECK("tanh_drop_bwd");
    return 0;
}

int hvae_rows_axpy(const int32_t* item, const int32_t* n_rows, int max_rows, const float* src, int ld_src, float alpha, float* dst,
                   int ld_dst, int cols, void* stream) {
    if (max_rows == 0) return 0;
    HVAE_REQUIRE(ld_src % 4 == 0 && ld_dst % 4 == 0 && cols % 4 == 0, "rows_axpy: leading dimensions and cols must be multiples of 4");
    launch_pdl(rows_axpy_kernel, max_rows, 128, 0, (cudaStream_t)stream, item, n_rows, max_rows, (const float4*)src, ld_src / 4, alpha,
               (float4*)dst, ld_dst / 4, cols / 4);
    HVAE_LAUNCH_CHECK("rows_axpy");
    return 0;
}

}  // extern "C"
