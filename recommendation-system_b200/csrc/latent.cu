// Latent-space kernels: reparameterisation + KL (reference src/ml/model.py:157-179,286-287), the
// GELU+dropout of the projection MLP (model.py:92-93), their backward, and the loss reduction
// (model.py:281-290).  All tiny and latency-bound; one warp per user row, fixed-order reductions.
#include "common.cuh"

namespace hvae {

// ml = [mu | logvar] rows of width 2L (leading dim ldml).  z = mu + eps*exp(0.5*logvar) (eps may be null: z = mu).
// kl_row[b] = -0.5 * sum_j (1 + lv - mu^2 - exp(lv)).
__global__ void reparam_kl_kernel(const float* __restrict__ ml, int ldml, const float* __restrict__ eps, int B, int L,
                                  float* __restrict__ z, int ldz, float* __restrict__ kl_row) {
    pdl_prologue();
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= B) return;
    const float* mu = ml + (size_t)row * ldml;
    const float* lv = mu + L;
    float s = 0.f;
    for (int j = lane; j < ldz; j += 32) {
        float zz = 0.f;
        if (j < L) {
            const float m = mu[j], v = lv[j];
            zz = eps ? m + eps[(size_t)row * L + j] * expf(0.5f * v) : m;
            s += 1.0f + v - m * m - expf(v);
        }
        z[(size_t)row * ldz + j] = zz;
    }
    s = warp_sum(s);
    if (lane == 0 && kl_row) kl_row[row] = -0.5f * s;
}

// dml = [dmu | dlv]:  dmu = dz + (beta/Bg)*mu ;  dlv = dz*eps*0.5*exp(0.5 lv) + (beta/Bg)*0.5*(exp(lv)-1)
// coef points at a device float holding beta / B_global.
__global__ void latent_bwd_kernel(const float* __restrict__ dz, int lddz, const float* __restrict__ ml, int ldml,
                                  const float* __restrict__ eps, int B, int L, const float* __restrict__ coef,
                                  float* __restrict__ dml) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * L) return;
    const int row = i / L, j = i - row * L;
    const float c = *coef;
    const float m = ml[(size_t)row * ldml + j], v = ml[(size_t)row * ldml + L + j];
    const float g = dz[(size_t)row * lddz + j];
    float dlv = c * 0.5f * (expf(v) - 1.0f);
    if (eps) dlv += g * eps[(size_t)row * L + j] * 0.5f * expf(0.5f * v);
    dml[(size_t)row * ldml + j] = g + c * m;
    dml[(size_t)row * ldml + L + j] = dlv;
}

// t = gelu(q) * mask * keep_scale  (mask may be null)
__global__ void gelu_drop_fwd_kernel(const float* __restrict__ q, const uint8_t* __restrict__ mask, float keep_scale, int B,
                                     int d, int ld, float* __restrict__ t) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * ld) return;
    const int row = i / ld, j = i - row * ld;
    float g = 0.f;
    if (j < d) {
        g = gelu(q[i]);
        if (mask) g = mask[(size_t)row * d + j] ? g * keep_scale : 0.f;
    }
    t[i] = g;
}

// dq = dt * mask * keep_scale * gelu'(q)   (dt may alias dq)
__global__ void gelu_drop_bwd_kernel(const float* dt, const float* __restrict__ q, const uint8_t* __restrict__ mask,
                                     float keep_scale, int B, int d, int ld, float* dq) {
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * ld) return;
    const int row = i / ld, j = i - row * ld;
    float g = 0.f;
    if (j < d) {
        g = dt[i];
        if (mask) g = mask[(size_t)row * d + j] ? g * keep_scale : 0.f;
        g *= gelu_grad(q[i]);
    }
    dq[i] = g;
}

// Single CTA, fixed order: recon = sum_b (xsum_b*lse_b - dot_b) / Bg ; kl = sum_b kl_row_b / Bg ;
// out = {recon + beta*kl, recon, kl}; acc[0..2] += out, acc[3] += 1  (per-epoch accumulators, read once).
__global__ void __launch_bounds__(1024) loss_finalize_kernel(const float* __restrict__ lse, const float* __restrict__ dot,
                                                              const float* __restrict__ xsum, const float* __restrict__ kl_row,
                                                              int B, const float* __restrict__ inv_bg,
                                                              const float* __restrict__ beta, float* __restrict__ out,
                                                              float* __restrict__ acc) {
    pdl_prologue();
    __shared__ double sr[32], sk[32];
    double r = 0.0, k = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        r += (double)xsum[b] * (double)lse[b] - (double)dot[b];
        k += (double)kl_row[b];
    }
    for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o);
        k += __shfl_xor_sync(0xffffffffu, k, o);
    }
    if ((threadIdx.x & 31) == 0) { sr[threadIdx.x >> 5] = r; sk[threadIdx.x >> 5] = k; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double R = 0.0, Kk = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { R += sr[w]; Kk += sk[w]; }
        const float recon = (float)(R * (double)*inv_bg), kl = (float)(Kk * (double)*inv_bg);
        const float total = recon + *beta * kl;
        out[0] = total; out[1] = recon; out[2] = kl;
        if (acc) { acc[0] += total; acc[1] += recon; acc[2] += kl; acc[3] += 1.0f; }
    }
}

}  // namespace hvae

using namespace hvae;

extern "C" {

int hvae_reparam_kl(const float* ml, int ldml, const float* eps, int B, int L, float* z, int ldz, float* kl_row, void* stream) {
    if (B == 0) return 0;
    launch_pdl(reparam_kl_kernel, ceil_div(B, 8), 256, 0, (cudaStream_t)stream, ml, ldml, eps, B, L, z, ldz, kl_row);
    HVAE_LAUNCH_CHECK("reparam_kl");
    return 0;
}

int hvae_latent_bwd(const float* dz, int lddz, const float* ml, int ldml, const float* eps, int B, int L, const float* coef,
                    float* dml, void* stream) {
    if (B == 0) return 0;
    launch_pdl(latent_bwd_kernel, ceil_div(B * L, 256), 256, 0, (cudaStream_t)stream, dz, lddz, ml, ldml, eps, B, L, coef, dml);
    HVAE_LAUNCH_CHECK("latent_bwd");
    return 0;
}

int hvae_gelu_drop_fwd(const float* q, const uint8_t* mask, float keep_scale, int B, int d, int ld, float* t, void* stream) {
    if (B == 0) return 0;
    launch_pdl(gelu_drop_fwd_kernel, ceil_div(B * ld, 256), 256, 0, (cudaStream_t)stream, q, mask, keep_scale, B, d, ld, t);
    HVAE_LAUNCH_CHECK("gelu_drop_fwd");
    return 0;
}

int hvae_gelu_drop_bwd(const float* dt, const float* q, const uint8_t* mask, float keep_scale, int B, int d, int ld, float* dq,
                       void* stream) {
    if (B == 0) return 0;
    launch_pdl(gelu_drop_bwd_kernel, ceil_div(B * ld, 256), 256, 0, (cudaStream_t)stream, dt, q, mask, keep_scale, B, d, ld, dq);
    HVAE_LAUNCH_CHECK("gelu_drop_bwd");
    return 0;
}

int hvae_loss_finalize(const float* lse, const float* dot, const float* xsum, const float* kl_row, int B, const float* inv_bg,
                       const float* beta, float* out, float* acc, void* stream) {
    launch_pdl(loss_finalize_kernel, 1, 1024, 0, (cudaStream_t)stream, lse, dot, xsum, kl_row, B, inv_bg, beta, out, acc);
    HVAE_LAUNCH_CHECK("loss_finalize");
    return 0;
}

}  // extern "C"
