// Encoder kernels.
//  (1) hvae_gather_ln_fwd : first encoder layer on the sparse user row as an embedding-bag gather-sum of
//      W1^T rows, fused with bias + LayerNorm + GELU + dropout.   Replaces nn.Linear(N,h) on a dense x
//      followed by LayerNorm/GELU/Dropout (reference src/ml/model.py:114-117,149).
//  (2) hvae_ln_act_fwd / hvae_ln_act_bwd : LayerNorm+GELU+dropout of the deeper hidden layers and the
//      backward of that block for every layer.
//  (3) hvae_w1_grad : backward of (1): per touched item, sum_b x_bi * dH_b, written to a compact
//      [n_unique, ld] gradient (deterministic, no atomics) together with its squared norm.
// HBM-bound; one warp owns one row, 128-bit loads, shuffle reductions.
#include <cstdlib>

#include "common.cuh"

namespace hvae {

__device__ __forceinline__ float& f4(float4& v, int k) { return reinterpret_cast<float*>(&v)[k]; }

template <int NCHUNK>
__device__ __forceinline__ void ln_gelu_drop_store(float4 (&acc)[NCHUNK], int lane, int row, int h, int ld4,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                   const uint8_t* __restrict__ mask, float keep_scale,
                                                   float* __restrict__ pre, float* __restrict__ mean_out,
                                                   float* __restrict__ rstd_out, float* __restrict__ act) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int col = (lane + 32 * c) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (col + k < h) s += f4(acc[c], k);
            else f4(acc[c], k) = 0.f;
        }
    }
    const float mean = warp_sum(s) / (float)h;
    float q = 0.f;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int col = (lane + 32 * c) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (col + k < h) { const float d = f4(acc[c], k) - mean; q += d * d; }
    }
    const float var = warp_sum(q) / (float)h;
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    if (lane == 0 && mean_out) { mean_out[row] = mean; rstd_out[row] = rstd; }
    const size_t base4 = (size_t)row * ld4;
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int col4 = lane + 32 * c;
        if (col4 >= ld4) continue;
        if (pre) reinterpret_cast<float4*>(pre)[base4 + col4] = acc[c];
        float4 o;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int col = col4 * 4 + k;
            float g = 0.f;
            if (col < h) {
                const float y = (f4(acc[c], k) - mean) * rstd * gamma[col] + beta[col];
                g = gelu(y);
                if (mask) g = mask[(size_t)row * h + col] ? g * keep_scale : 0.f;
            }
            f4(o, k) = g;
        }
        reinterpret_cast<float4*>(act)[base4 + col4] = o;
    }
}

template <int NCHUNK>
__global__ void __launch_bounds__(256) gather_ln_fwd_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
    const int32_t* __restrict__ rows, int B, const float4* __restrict__ W1T, int ld4, int h,
    const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
    const uint8_t* __restrict__ mask, float keep_scale, float* __restrict__ pre, float* __restrict__ mean_out,
    float* __restrict__ rstd_out, float* __restrict__ act) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int u = rows ? rows[warp] : warp;
    const int64_t s = indptr[u], e = indptr[u + 1];
    float4 acc[NCHUNK];
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int64_t j0 = s; j0 < e; j0 += 32) {
        const int cnt = (int)min((int64_t)32, e - j0);
        int my_idx = 0;
        float my_val = 0.f;
        if (lane < cnt) {
            my_idx = indices[j0 + lane];
            my_val = values ? values[j0 + lane] : 1.0f;
        }
        int t = 0;
        for (; t + 2 <= cnt; t += 2) {  // two rows in flight per iteration
            const int i0 = __shfl_sync(0xffffffffu, my_idx, t), i1 = __shfl_sync(0xffffffffu, my_idx, t + 1);
            const float v0 = __shfl_sync(0xffffffffu, my_val, t), v1 = __shfl_sync(0xffffffffu, my_val, t + 1);
            const float4* r0 = W1T + (size_t)i0 * ld4;
            const float4* r1 = W1T + (size_t)i1 * ld4;
            float4 a[NCHUNK], b[NCHUNK];
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                const int col4 = lane + 32 * c;
                if (col4 < ld4) { a[c] = __ldg(r0 + col4); b[c] = __ldg(r1 + col4); }
                else { a[c] = make_float4(0.f, 0.f, 0.f, 0.f); b[c] = a[c]; }
            }
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
#pragma unroll
                for (int k = 0; k < 4; ++k) f4(acc[c], k) = fmaf(v0, f4(a[c], k), f4(acc[c], k));
#pragma unroll
                for (int k = 0; k < 4; ++k) f4(acc[c], k) = fmaf(v1, f4(b[c], k), f4(acc[c], k));
            }
        }
        if (t < cnt) {
            const int i0 = __shfl_sync(0xffffffffu, my_idx, t);
            const float v0 = __shfl_sync(0xffffffffu, my_val, t);
            const float4* r0 = W1T + (size_t)i0 * ld4;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                const int col4 = lane + 32 * c;
                if (col4 < ld4) {
                    const float4 a = __ldg(r0 + col4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) f4(acc[c], k) = fmaf(v0, reinterpret_cast<const float*>(&a)[k], f4(acc[c], k));
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int col = (lane + 32 * c) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (col + k < h) f4(acc[c], k) += bias[col + k];
    }
    if (gamma) {
        ln_gelu_drop_store<NCHUNK>(acc, lane, warp, h, ld4, gamma, beta, mask, keep_scale, pre, mean_out, rstd_out, act);
    } else {  // plain linear output (used by tests of the gather alone)
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            const int col4 = lane + 32 * c;
            if (col4 < ld4) reinterpret_cast<float4*>(act)[(size_t)warp * ld4 + col4] = acc[c];
        }
    }
}

// Small batches (a few hundred users) leave a warp-per-user grid far below the machine's latency-hiding capacity:
// this variant gives every user a whole CTA, one thread per 16-byte column chunk, so ~5x more row loads are in
// flight.  Same arithmetic per element (entries accumulated in CSR order); LayerNorm statistics via a block reduction.
__global__ void __launch_bounds__(512) gather_ln_fwd_block_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
    const int32_t* __restrict__ rows, int B, const float4* __restrict__ W1T, int ld4, int h,
    const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
    const uint8_t* __restrict__ mask, float keep_scale, float* __restrict__ pre, float* __restrict__ mean_out,
    float* __restrict__ rstd_out, float* __restrict__ act) {
    pdl_prologue();
    __shared__ float stat[2];
    const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int u = rows ? rows[b] : b;
    const int64_t s = indptr[u], e = indptr[u + 1];
    const bool on = t < ld4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // eight rows in flight per thread; the ragged tail is a predicated batch (not a serial loop: every dependent
    // indices -> row round trip costs a full memory latency), entries still accumulated in CSR order
    for (int64_t j = s; j < e; j += 8) {
        const int n = (int)min((int64_t)8, e - j);
        int id[8]; float xv[8]; float4 w[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            id[q] = q < n ? indices[j + q] : 0;
            xv[q] = (values && q < n) ? values[j + q] : 1.0f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) w[q] = (on && q < n) ? __ldg(W1T + (size_t)id[q] * ld4 + t) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q < n) {
                acc.x = fmaf(xv[q], w[q].x, acc.x); acc.y = fmaf(xv[q], w[q].y, acc.y);
                acc.z = fmaf(xv[q], w[q].z, acc.z); acc.w = fmaf(xv[q], w[q].w, acc.w);
            }
        }
    }
    const int col = t * 4;
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (on && col + k < h) v[k] += bias[col + k];
        else v[k] = 0.f;
    }
    if (!gamma) {   // plain linear output
        if (on) reinterpret_cast<float4*>(act)[(size_t)b * ld4 + t] = make_float4(v[0], v[1], v[2], v[3]);
        return;
    }
    // LayerNorm statistics in exactly the summation order of the warp-per-user kernel (lane l adds its columns
    // (l+32c)*4+k sequentially over c, k; then the xor-shuffle tree), so both variants give bit-identical results.
    extern __shared__ float sv[];   // [ld4 * 4]
    if (on) reinterpret_cast<float4*>(sv)[t] = make_float4(v[0], v[1], v[2], v[3]);
    __syncthreads();
    if (warp == 0) {
        float sm = 0.f;
        for (int c4 = lane; c4 < ld4; c4 += 32)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c4 * 4 + k < h) sm += sv[c4 * 4 + k];
        const float mean_w = warp_sum(sm) / (float)h;
        float q = 0.f;
        for (int c4 = lane; c4 < ld4; c4 += 32)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (c4 * 4 + k < h) { const float d = sv[c4 * 4 + k] - mean_w; q += d * d; }
        const float var = warp_sum(q) / (float)h;
        if (lane == 0) {
            stat[0] = mean_w;
            stat[1] = 1.0f / sqrtf(var + 1e-5f);
            if (mean_out) { mean_out[b] = mean_w; rstd_out[b] = stat[1]; }
        }
    }
    __syncthreads();
    const float mean = stat[0];
    const float rstd = stat[1];
    if (!on) return;
    if (pre) reinterpret_cast<float4*>(pre)[(size_t)b * ld4 + t] = make_float4(v[0], v[1], v[2], v[3]);
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float g = 0.f;
        if (col + k < h) {
            const float y = (v[k] - mean) * rstd * gamma[col + k] + beta[col + k];
            g = gelu(y);
            if (mask) g = mask[(size_t)b * h + col + k] ? g * keep_scale : 0.f;
        }
        o[k] = g;
    }
    reinterpret_cast<float4*>(act)[(size_t)b * ld4 + t] = make_float4(o[0], o[1], o[2], o[3]);
}

template <int NCHUNK>
__global__ void __launch_bounds__(256) ln_act_fwd_kernel(const float* __restrict__ pre, int B, int h, int ld4,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const uint8_t* __restrict__ mask, float keep_scale,
                                                         float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                         float* __restrict__ act) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    float4 acc[NCHUNK];
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) {
        const int col4 = lane + 32 * c;
        acc[c] = col4 < ld4 ? reinterpret_cast<const float4*>(pre)[(size_t)warp * ld4 + col4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ln_gelu_drop_store<NCHUNK>(acc, lane, warp, h, ld4, gamma, beta, mask, keep_scale, nullptr, mean_out, rstd_out, act);
}

// Backward of LayerNorm+GELU+dropout.  dact may alias dpre.  Each warp walks rows warp, warp+W, ... and
// keeps its share of d(gamma), d(beta) in registers; per-warp partials go to `partial` [nwarps][2][ld]
// and are summed in a fixed order by colsum (deterministic).
template <int NCHUNK>
__global__ void __launch_bounds__(256) ln_act_bwd_kernel(const float* dact, const float* __restrict__ pre,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const uint8_t* __restrict__ mask, float keep_scale, int B, int h,
                                                         int ld4, float* dpre, float* __restrict__ partial) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    float4 ag[NCHUNK], ab[NCHUNK];
#pragma unroll
    for (int c = 0; c < NCHUNK; ++c) ag[c] = ab[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int row = warp; row < B; row += nwarps) {
        const float mu = mean[row], rs = rstd[row];
        float4 xh[NCHUNK], dx[NCHUNK];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            const int col4 = lane + 32 * c;
            xh[c] = dx[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col4 >= ld4) continue;
            const float4 p = reinterpret_cast<const float4*>(pre)[(size_t)row * ld4 + col4];
            const float4 da = reinterpret_cast<const float4*>(dact)[(size_t)row * ld4 + col4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int col = col4 * 4 + k;
                if (col >= h) continue;
                const float xhat = (reinterpret_cast<const float*>(&p)[k] - mu) * rs;
                const float y = xhat * gamma[col] + beta[col];
                float dg = reinterpret_cast<const float*>(&da)[k];
                if (mask) dg = mask[(size_t)row * h + col] ? dg * keep_scale : 0.f;
                const float dln = dg * gelu_grad(y);
                f4(ag[c], k) += dln * xhat;
                f4(ab[c], k) += dln;
                const float dxh = dln * gamma[col];
                f4(xh[c], k) = xhat;
                f4(dx[c], k) = dxh;
                s1 += dxh;
                s2 += dxh * xhat;
            }
        }
        const float m1 = warp_sum(s1) / (float)h, m2 = warp_sum(s2) / (float)h;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            const int col4 = lane + 32 * c;
            if (col4 >= ld4) continue;
            float4 o;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int col = col4 * 4 + k;
                f4(o, k) = col < h ? rs * (f4(dx[c], k) - m1 - f4(xh[c], k) * m2) : 0.f;
            }
            reinterpret_cast<float4*>(dpre)[(size_t)row * ld4 + col4] = o;
        }
    }
    // block-level reduction in a fixed warp order (deterministic), one partial row [2][ld] per block
    extern __shared__ float4 bred[];  // [2][ld4]
    const int wib = threadIdx.x >> 5;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
        if (wib == w) {
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                const int col4 = lane + 32 * c;
                if (col4 >= ld4) continue;
                if (w == 0) { bred[col4] = ag[c]; bred[ld4 + col4] = ab[c]; }
                else {
                    float4 a = bred[col4], b = bred[ld4 + col4];
                    a.x += ag[c].x; a.y += ag[c].y; a.z += ag[c].z; a.w += ag[c].w;
                    b.x += ab[c].x; b.y += ab[c].y; b.z += ab[c].z; b.w += ab[c].w;
                    bred[col4] = a; bred[ld4 + col4] = b;
                }
            }
        }
        __syncthreads();
    }
    float4* pg = reinterpret_cast<float4*>(partial) + (size_t)blockIdx.x * 2 * ld4;
    for (int i = threadIdx.x; i < 2 * ld4; i += blockDim.x) pg[i] = bred[i];
}

// out[chunk][c] = sum over rows r in [chunk*rpc, min(R,(chunk+1)*rpc)) of X[r][c]   (fixed order)
__global__ void colsum_chunk_kernel(const float* __restrict__ X, int ld, int R, int C, int rpc, float* __restrict__ out,
                                    int out_ld) {
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int r0 = blockIdx.y * rpc, r1 = min(R, r0 + rpc);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f, s5 = 0.f, s6 = 0.f, s7 = 0.f;
    int r = r0;
    for (; r + 8 <= r1; r += 8) {
        s0 += X[(size_t)r * ld + c];
        s1 += X[(size_t)(r + 1) * ld + c];
        s2 += X[(size_t)(r + 2) * ld + c];
        s3 += X[(size_t)(r + 3) * ld + c];
        s4 += X[(size_t)(r + 4) * ld + c];
        s5 += X[(size_t)(r + 5) * ld + c];
        s6 += X[(size_t)(r + 6) * ld + c];
        s7 += X[(size_t)(r + 7) * ld + c];
    }
    for (; r < r1; ++r) s0 += X[(size_t)r * ld + c];
    out[(size_t)blockIdx.y * out_ld + c] = ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
}

// Layer-1 weight gradient.  A touched item's gradient row = sum over its (user, value) entries of value * dH[user, :].
// Item popularity is heavy-tailed (the top item of a Zipf catalogue sits in most rows of a batch), so the work is cut
// into items of at most kW1Chunk entries (hvae_w1_plan, run beside the forward pass): one CTA (4 warps) per work item.
// Entries were sorted stably by item, so every sum has a fixed order: warp w adds entries w, w+4, ... of its chunk, the
// four warp sums combine as (0+1)+(2+3); an item with several chunks gets its partial rows added in chunk order by
// w1_combine_kernel.  Entry ids are fetched 32 at a time (one per lane) and broadcast, every row is read as NCHUNK
// 16-byte vectors per lane with two rows in flight.
constexpr int kW1Chunk = 64;

// Single CTA: chunk_base[s] = first work item of slot s, part_base[s] = first partial row of slot s (slots with one chunk
// have none), n_work = chunk_base[n_unique].
__global__ void __launch_bounds__(1024) w1_plan_kernel(const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_unique,
                                                        int32_t* __restrict__ chunk_base, int32_t* __restrict__ part_base,
                                                        int32_t* __restrict__ n_work) {
    pdl_prologue();
    __shared__ int wsum_a[32], wsum_b[32];
    __shared__ int carry_a, carry_b;
    const int n = *n_unique, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; }
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int sl = base + threadIdx.x;
        int a = 0, b = 0;
        if (sl < n) { a = (seg_start[sl + 1] - seg_start[sl] + kW1Chunk - 1) / kW1Chunk; b = a > 1 ? a : 0; }
        int ia = a, ib = b;                      // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int ta = __shfl_up_sync(0xffffffffu, ia, o), tb2 = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += ta; ib += tb2; }
        }
        if (lane == 31) { wsum_a[warp] = ia; wsum_b[warp] = ib; }
        __syncthreads();
        if (warp == 0) {
            int va = wsum_a[lane], vb = wsum_b[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ta = __shfl_up_sync(0xffffffffu, va, o), tb2 = __shfl_up_sync(0xffffffffu, vb, o);
                if (lane >= o) { va += ta; vb += tb2; }
            }
            wsum_a[lane] = va; wsum_b[lane] = vb;
        }
        __syncthreads();
        const int offa = carry_a + (warp ? wsum_a[warp - 1] : 0), offb = carry_b + (warp ? wsum_b[warp - 1] : 0);
        if (sl < n) { chunk_base[sl] = offa + ia - a; part_base[sl] = offb + ib - b; }
        __syncthreads();
        if (threadIdx.x == 0) { carry_a += wsum_a[31]; carry_b += wsum_b[31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { chunk_base[n] = carry_a; part_base[n] = carry_b; n_work[0] = carry_a; n_work[1] = 0; }
}

// work_slot[w] = slot of work item w; multi_slot[] = the slots that have several chunks (order irrelevant: every entry is
// combined independently), *n_multi their count (reset by the plan kernel).
__global__ void w1_expand_kernel(const int32_t* __restrict__ n_unique, const int32_t* __restrict__ chunk_base,
                                 int32_t* __restrict__ work_slot, int max_slots, int32_t* __restrict__ multi_slot,
                                 int32_t* __restrict__ n_multi) {
    pdl_prologue();
    const int sl = blockIdx.x * blockDim.x + threadIdx.x;
    if (sl >= max_slots || sl >= *n_unique) return;
    const int w0 = chunk_base[sl], w1 = chunk_base[sl + 1];
    for (int w = w0; w < w1; ++w) work_slot[w] = sl;
    if (w1 - w0 > 1) multi_slot[atomicAdd(n_multi, 1)] = sl;
}
// One WARP per work item, a persistent grid striding over the items: most touched items of a Zipf catalogue have one or two
// entries, so a CTA per item (with a shared-memory reduction over its warps, and a grid sized for the worst case whose CTAs
// mostly exit at once) spent its time on launch overhead and barriers -- 0.55 ms at the 8-GPU global batch.  The warp adds the
// item's entries in their (stable, item-sorted) order, up to four dH rows in flight; the sum order is fixed, hence deterministic.
template <int NCHUNK>
__global__ void __launch_bounds__(256) w1_grad_kernel(const int32_t* __restrict__ seg_start, const int32_t* __restrict__ n_work,
                                                      const int32_t* __restrict__ work_slot, const int32_t* __restrict__ chunk_base,
                                                      const int32_t* __restrict__ part_base,
                                                      const int32_t* __restrict__ sorted_eid, const int32_t* __restrict__ ent_user,
                                                      const float* __restrict__ ent_val, const float* __restrict__ dpre, int ld4,
                                                      int block_rows, int64_t block_stride4, float* __restrict__ gs,
                                                      float* __restrict__ partial, float* __restrict__ rownorm2) {
    pdl_prologue();
    // dpre row of batch position u: blocks of `block_rows` rows, `block_stride4` float4 apart (data-parallel training reads
    // the all-gathered per-rank buffers in place); a plain [rows, ld] matrix has block_rows = INT_MAX.
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const int nw = *n_work;
    const float4* dp = reinterpret_cast<const float4*>(dpre);
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nw; w += warps) {
        const int slot = work_slot[w];
        const int cb = chunk_base[slot], nchunks = chunk_base[slot + 1] - cb, chunk = w - cb;
        const int s = seg_start[slot] + chunk * kW1Chunk, e = min(seg_start[slot + 1], s + kW1Chunk);
        float* out_row = nchunks == 1 ? gs + (size_t)slot * ld4 * 4 : partial + (size_t)(part_base[slot] + chunk) * ld4 * 4;
        float4 a[NCHUNK];
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) a[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = s; base < e; base += 32) {
            const int cnt = min(32, e - base);
            int my_user = 0;
            float my_val = 0.f;
            if (lane < cnt) { const int eid = sorted_eid[base + lane]; my_user = ent_user[eid]; my_val = ent_val[eid]; }
            constexpr int R = NCHUNK <= 6 ? 4 : NCHUNK <= 8 ? 2 : 1;        // dH rows in flight (register budget)
            for (int t = 0; t < cnt; t += R) {
                float4 g[R][NCHUNK];
                float v[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int tt = min(t + r, cnt - 1);                  // clamped; its weight is zeroed below
                    const int u = __shfl_sync(0xffffffffu, my_user, tt);
                    v[r] = t + r < cnt ? __shfl_sync(0xffffffffu, my_val, tt) : 0.f;
                    const size_t o = (size_t)(u / block_rows) * block_stride4 + (size_t)(u % block_rows) * ld4;
#pragma unroll
                    for (int c = 0; c < NCHUNK; ++c) {
                        const int col4 = lane + 32 * c;
                        g[r][c] = col4 < ld4 ? __ldg(dp + o + col4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (t + r >= cnt) break;                             // (a zero weight would still turn an inf into a NaN)
#pragma unroll
                    for (int c = 0; c < NCHUNK; ++c) {
                        a[c].x = fmaf(v[r], g[r][c].x, a[c].x); a[c].y = fmaf(v[r], g[r][c].y, a[c].y);
                        a[c].z = fmaf(v[r], g[r][c].z, a[c].z); a[c].w = fmaf(v[r], g[r][c].w, a[c].w);
                    }
                }
            }
        }
        float n2 = 0.f;
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c) {
            const int col4 = lane + 32 * c;
            if (col4 < ld4) {
                reinterpret_cast<float4*>(out_row)[col4] = a[c];
                n2 += a[c].x * a[c].x + a[c].y * a[c].y + a[c].z * a[c].z + a[c].w * a[c].w;
            }
        }
        if (nchunks == 1) {            // (the norm of a multi-chunk row is taken after the combine)
            n2 = warp_sum(n2);
            if (lane == 0) rownorm2[slot] = n2;
        }
    }
}

// Items with several chunks: gs[slot] = partial rows added in chunk order; row norm.  One CTA per multi-chunk slot.
__global__ void __launch_bounds__(128) w1_combine_kernel(const int32_t* __restrict__ n_multi, const int32_t* __restrict__ multi_slot,
                                                         const int32_t* __restrict__ chunk_base,
                                                         const int32_t* __restrict__ part_base, const float* __restrict__ partial,
                                                         int ld4, float* __restrict__ gs, float* __restrict__ rownorm2) {
    pdl_prologue();
    if ((int)blockIdx.x >= *n_multi) return;
    const int slot = multi_slot[blockIdx.x];
    const int nchunks = chunk_base[slot + 1] - chunk_base[slot];
    const float4* pr = reinterpret_cast<const float4*>(partial) + (size_t)part_base[slot] * ld4;
    float n2 = 0.f;
    for (int col4 = threadIdx.x; col4 < ld4; col4 += 128) {
        float4 a = pr[col4];
        for (int c = 1; c < nchunks; ++c) {
            const float4 b = pr[(size_t)c * ld4 + col4];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        reinterpret_cast<float4*>(gs)[(size_t)slot * ld4 + col4] = a;
        n2 += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    }
    __shared__ float wsum[4];
    n2 = warp_sum(n2);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = n2;
    __syncthreads();
    if (threadIdx.x == 0) rownorm2[slot] = (wsum[0] + wsum[1]) + (wsum[2] + wsum[3]);
}

// dense (autograd-compat) variant of the layer-1 weight gradient: dW1T[item, :] += x * dH[b, :]
__global__ void w1_grad_dense_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                     const float* __restrict__ values, const int32_t* __restrict__ rows, int B,
                                     const float* __restrict__ dpre, int ld, int h, float* __restrict__ dW1T) {
    pdl_prologue();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= B) return;
    const int u = rows ? rows[warp] : warp;
    for (int64_t j = indptr[u]; j < indptr[u + 1]; ++j) {
        const float v = values ? values[j] : 1.0f;
        float* dst = dW1T + (size_t)indices[j] * ld;
        for (int c = lane; c < h; c += 32) atomicAdd(dst + c, v * dpre[(size_t)warp * ld + c]);
    }
}

}  // namespace hvae

using namespace hvae;

#define DISPATCH_NCHUNK(nch, ...)                                       \
    switch (nch) {                                                      \
        case 1: { constexpr int NC = 1; __VA_ARGS__; } break;           \
        case 2: { constexpr int NC = 2; __VA_ARGS__; } break;           \
        case 3: { constexpr int NC = 3; __VA_ARGS__; } break;           \
        case 4: { constexpr int NC = 4; __VA_ARGS__; } break;           \
        case 5: { constexpr int NC = 5; __VA_ARGS__; } break;           \
        case 6: { constexpr int NC = 6; __VA_ARGS__; } break;           \
        case 7: case 8: { constexpr int NC = 8; __VA_ARGS__; } break;   \
        case 9: case 10: case 11: case 12: { constexpr int NC = 12; __VA_ARGS__; } break; \
        case 13: case 14: case 15: case 16: { constexpr int NC = 16; __VA_ARGS__; } break; \
        default: return hvae_fail("hidden width %d too large (max 2048)", h); \
    }

extern "C" {

int hvae_gather_ln_fwd(const int64_t* indptr, const int32_t* indices, const float* values, const int32_t* rows, int B,
                       const float* W1T, int ld, int h, const float* bias, const float* gamma, const float* beta,
                       const uint8_t* mask, float keep_scale, float* pre, float* mean, float* rstd, float* act,
                       void* stream) {
    HVAE_REQUIRE(ld % 4 == 0 && ld >= h, "gather_ln_fwd: ld=%d must be a multiple of 4 and >= h=%d", ld, h);
    if (B == 0) return 0;
    if (ld / 4 <= 512 && getenv("HVAE_GATHER_WARP") == nullptr) {   // one CTA per user (more loads in flight than a warp per user)
        launch_pdl(gather_ln_fwd_block_kernel, B, round_up(ld / 4, 32), (size_t)ld * sizeof(float), (cudaStream_t)stream, 
            indptr, indices, values, rows, B, reinterpret_cast<const float4*>(W1T), ld / 4, h, bias, gamma, beta, mask, keep_scale, pre,
            mean, rstd, act);
        HVAE_LAUNCH_CHECK("gather_ln_fwd(block)");
        return 0;
    }
    const int nch = ceil_div(ld / 4, 32);
    const int blocks = ceil_div(B, 8);
    DISPATCH_NCHUNK(nch, (launch_pdl(gather_ln_fwd_kernel<NC>, blocks, 256, 0, (cudaStream_t)stream, 
                             indptr, indices, values, rows, B, reinterpret_cast<const float4*>(W1T), ld / 4, h, bias,
                             gamma, beta, mask, keep_scale, pre, mean, rstd, act)));
    HVAE_LAUNCH_CHECK("gather_ln_fwd");
    return 0;
}

int hvae_ln_act_fwd(const float* pre, int B, int h, int ld, const float* gamma, const float* beta, const uint8_t* mask,
                    float keep_scale, float* mean, float* rstd, float* act, void* stream) {
    HVAE_REQUIRE(ld % 4 == 0 && ld >= h, "ln_act_fwd: bad ld=%d h=%d", ld, h);
    if (B == 0) return 0;
    const int nch = ceil_div(ld / 4, 32);
    DISPATCH_NCHUNK(nch, (launch_pdl(ln_act_fwd_kernel<NC>, ceil_div(B, 8), 256, 0, (cudaStream_t)stream, 
                             pre, B, h, ld / 4, gamma, beta, mask, keep_scale, mean, rstd, act)));
    HVAE_LAUNCH_CHECK("ln_act_fwd");
    return 0;
}

// workspace: partial [nwarps = 8*blocks][2][ld] floats + stage buffer; see hvae_ln_bwd_workspace_floats
// one row per warp and trip; up to two 8-warp blocks per SM (a 4,096-row batch ran 43 us on 74 blocks: 7 dependent row trips per warp)
static inline int ln_bwd_blocks(int B) { return max(1, min(2 * kNumSMs, ceil_div(B, 8))); }
constexpr int kLnBwdOnePass = 74;      // up to this many partial rows are summed by one launch, more in two stages

size_t hvae_ln_bwd_workspace_floats(int B, int ld) { return ((size_t)ln_bwd_blocks(B) * 8 + 64) * 2 * ld; }

int hvae_ln_act_bwd(const float* dact, const float* pre, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, const uint8_t* mask, float keep_scale, int B, int h, int ld, float* dpre,
                    float* dgamma, float* dbeta, float* workspace, void* stream) {
    HVAE_REQUIRE(ld % 4 == 0 && ld >= h, "ln_act_bwd: bad ld=%d h=%d", ld, h);
    if (B == 0) return 0;
    const int nch = ceil_div(ld / 4, 32);
    const int blocks = ln_bwd_blocks(B);
    const size_t smem = (size_t)2 * ld * sizeof(float);
    HVAE_REQUIRE(smem <= 48 * 1024, "ln_act_bwd: hidden width %d too large", ld);
    DISPATCH_NCHUNK(nch, (launch_pdl(ln_act_bwd_kernel<NC>, blocks, 256, smem, (cudaStream_t)stream, 
                             dact, pre, mean, rstd, gamma, beta, mask, keep_scale, B, h, ld / 4, dpre, workspace)));
    HVAE_LAUNCH_CHECK("ln_act_bwd");
    // one partial row per block, [2*ld] wide: first ld = d(gamma), next ld = d(beta)
    int R = blocks;
    if (R > kLnBwdOnePass) {      // first stage: 64 groups of partial rows -> 64 rows behind the partials
        const int chunks = 64, rpc = ceil_div(R, chunks);
        float* stage = workspace + (size_t)blocks * 8 * 2 * ld;
        launch_pdl(colsum_chunk_kernel, dim3(ceil_div(2 * ld, 128), chunks), 128, 0, (cudaStream_t)stream, workspace, 2 * ld, R, 2 * ld, rpc, stage, 2 * ld);
        workspace = stage;
        R = ceil_div(R, rpc);
    }
    if (dbeta == dgamma + ld) {   // adjacent slots of the gradient arena: one launch over both (pad columns of the partials are zero)
        launch_pdl(colsum_chunk_kernel, dim3(ceil_div(2 * ld, 64), 1), 64, 0, (cudaStream_t)stream, workspace, 2 * ld, R, 2 * ld, R, dgamma, 2 * ld);
    } else {
        launch_pdl(colsum_chunk_kernel, dim3(ceil_div(h, 128), 1), 128, 0, (cudaStream_t)stream, workspace, 2 * ld, R, h, R, dgamma, h);
        launch_pdl(colsum_chunk_kernel, dim3(ceil_div(h, 128), 1), 128, 0, (cudaStream_t)stream, workspace + ld, 2 * ld, R, h, R, dbeta, h);
    }
    HVAE_LAUNCH_CHECK("ln_act_bwd colsum");
    return 0;
}

// out[c] = sum_r X[r][c]; two fixed-order stages through `workspace` (>= 64*C floats)
int hvae_colsum(const float* X, int ld, int R, int C, float* out, float* workspace, void* stream) {
    if (C == 0) return 0;
    const int chunks = max(1, min(64, ceil_div(R, 16)));
    const int rpc = ceil_div(max(R, 1), chunks);
    if (chunks == 1) {   // few rows (e.g. the per-rank partial gradients of data-parallel training): one pass
        launch_pdl(colsum_chunk_kernel, dim3(ceil_div(C, 128), 1), 128, 0, (cudaStream_t)stream, X, ld, R, C, R, out, C);
        HVAE_LAUNCH_CHECK("colsum");
        return 0;
    }
    launch_pdl(colsum_chunk_kernel, dim3(ceil_div(C, 128), chunks), 128, 0, (cudaStream_t)stream, X, ld, R, C, rpc, workspace, C);
    launch_pdl(colsum_chunk_kernel, dim3(ceil_div(C, 128), 1), 128, 0, (cudaStream_t)stream, workspace, C, chunks, C, chunks, out, C);
    HVAE_LAUNCH_CHECK("colsum");
    return 0;
}

// Work plan of hvae_w1_grad for the transposed batch (depends on the batch only).  chunk_base, part_base: int32 [max_slots+1];
// work_slot: int32 [hvae_w1_max_work(max_slots)]; n_work: int32 [2] = {work items, multi-chunk slots};
// multi_slot: int32 [hvae_w1_max_partial_rows(max_slots)].
size_t hvae_w1_max_work(int max_slots) { return (size_t)max_slots + (size_t)max_slots / kW1Chunk + 1; }
size_t hvae_w1_max_partial_rows(int max_slots) { return 2 * ((size_t)max_slots / kW1Chunk + 1); }

int hvae_w1_plan(const int32_t* seg_start, const int32_t* n_unique, int max_slots, int32_t* chunk_base, int32_t* part_base,
                 int32_t* work_slot, int32_t* multi_slot, int32_t* n_work, void* stream) {
    if (max_slots == 0) return 0;
    launch_pdl(w1_plan_kernel, 1, 1024, 0, (cudaStream_t)stream, seg_start, n_unique, chunk_base, part_base, n_work);
    launch_pdl(w1_expand_kernel, ceil_div(max_slots, 256), 256, 0, (cudaStream_t)stream, n_unique, chunk_base, work_slot, max_slots, multi_slot,
               n_work + 1);
    HVAE_LAUNCH_CHECK("w1_plan");
    return 0;
}

int hvae_w1_grad(const int32_t* seg_start, const int32_t* n_unique, const int32_t* sorted_eid, const int32_t* ent_user,
                 const float* ent_val, int max_slots, const int32_t* chunk_base, const int32_t* part_base, const int32_t* work_slot,
                 const int32_t* multi_slot, const int32_t* n_work, const float* dpre, int ld, int block_rows, int64_t block_stride,
                 float* gs, float* partial, float* rownorm2, void* stream) {
    HVAE_REQUIRE(ld % 4 == 0 && block_stride % 4 == 0, "w1_grad: ld=%d and the block stride must be multiples of 4", ld);
    if (block_rows <= 0) { block_rows = 0x7fffffff; block_stride = 0; }
    if (max_slots == 0) return 0;
    const int h = ld;
    const int nch = ceil_div(ld / 4, 32);
    const int max_work = (int)hvae_w1_max_work(max_slots);
    const int grid = max(1, min(kNumSMs * 8, ceil_div(max_work, 8)));      // persistent: 8 warps per CTA stride over the work items
    DISPATCH_NCHUNK(nch, (launch_pdl(w1_grad_kernel<NC>, grid, 256, 0, (cudaStream_t)stream, 
                             seg_start, n_work, work_slot, chunk_base, part_base, sorted_eid, ent_user, ent_val, dpre, ld / 4, block_rows,
                             block_stride / 4, gs, partial, rownorm2)));
    launch_pdl(w1_combine_kernel, (int)hvae_w1_max_partial_rows(max_slots), 128, 0, (cudaStream_t)stream, n_work + 1, multi_slot, chunk_base,
               part_base, partial, ld / 4, gs, rownorm2);
    HVAE_LAUNCH_CHECK("w1_grad");
    return 0;
}

int hvae_w1_grad_dense(const int64_t* indptr, const int32_t* indices, const float* values, const int32_t* rows, int B,
                       const float* dpre, int ld, int h, float* dW1T, void* stream) {
    if (B == 0) return 0;
    launch_pdl(w1_grad_dense_kernel, ceil_div(B, 8), 256, 0, (cudaStream_t)stream, indptr, indices, values, rows, B, dpre, ld, h, dW1T);
    HVAE_LAUNCH_CHECK("w1_grad_dense");
    return 0;
}

}  // extern "C"
