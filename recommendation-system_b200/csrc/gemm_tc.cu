// Small dense GEMMs of the MLP stack on the tensor cores (bf16/"tensor-core" mode of the model):
//   C[m,n] = alpha * sum_k A(m,k) * B(k,n) (+ bias[n]),   fp32 operands read as TF32, fp32 accumulation in TMEM.
// Same contract as hvae_gemm_f32 (arbitrary transposes through strides), so the engine swaps one for the other:
// fc_mu/fc_logvar, the projection MLP and the deeper hidden layers, forward, dX and dW
// (reference src/ml/model.py:90-95,114,126-127 and their autograd).  At a 512-user batch these nine GEMMs are
// 400-600 wide and latency-bound; a TMA-fed tcgen05 tile (128 x 64, k-blocks of 32 floats = one 128B swizzle span)
// turns each into a few microseconds.  The fp32 ("exact") mode keeps the FFMA kernel.
// The MMA reads both operands K-major (128B-swizzled rows of 32 floats).  An operand whose m / n axis is the
// contiguous one in memory (the transposed uses: dX = dY W, dW = dY^T X) is TMA-loaded un-swizzled into a staging
// area and transposed smem -> smem by the eight conversion warps (which also round every operand to TF32 and later run the
// epilogue; kind::tf32 does not accept MN-major shared-memory descriptors in the 128B-swizzle form: measured, it yields zeros).
// Epilogue: each 32 x 32 accumulator chunk goes TMEM -> registers -> a per-warp shared-memory tile -> coalesced global stores;
// optional fused epilogues (TG_EPI_*) and bias-gradient column sums, see include/hvae_b200.h.
#include <cstdlib>

#include "common.cuh"
#include "hvae_b200.h"
#include "tc_common.cuh"

namespace hvae {
namespace tc {

// Output tile 128 x BN with BN = 64 (small batches: more CTAs, split-K over a cluster), 128 or 256 (large batches: the TF32 rounding
// pass over the A tile is amortised over 2-4x more MMA work, one wave of CTAs instead of three).
constexpr int TG_BM = 128, TG_BK = 32, TG_MAX_STAGES = 8;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 4;
// stage = [A tile | B tile | A raw (only if A is MN-major) | B raw (only if B is MN-major)]; as many stages as fit (<= 8)
constexpr int TG_SMEM_BUDGET = 196608;
constexpr int TG_CONV_WARPS = 8, TG_THREADS = 64 + 32 * TG_CONV_WARPS;

struct TGemmParams {
    int M, N, K;
    int a_mn, b_mn;   // 1: operand is MN-major (m resp. n contiguous in memory)
    int stage_bytes, n_stages;
    int conv_per_stage;   // conversion warps that share one ring stage (8 warps over n_stages stages)
    int ksplit;       // CTAs per output tile (cluster along z): each takes a contiguous range of k-blocks
    int direct;       // debug (HVAE_TF32_TRUNC=1, both operands K-major): skip the rounding pass, the MMA truncates
    float* C;
    int64_t ldc;
    const float* bias;
    float alpha;
    // ---- fused epilogue (TG_EPI_*): the element-wise kernel that used to follow this GEMM, and the bias-gradient column sums
    int epi;
    int n_store;                        // columns written per row (>= N: pad columns up to the leading dimension get zeros)
    float* C2;                          // GELU_DROP: act = gelu(C) * mask * keep_scale (leading dimension ldc)
    __nv_bfloat16* Cb; int64_t ldcb;    // BF16: bf16 copy of C (C itself may be null)
    const float* aux; int64_t ldaux;    // GELU_BWD: pre-activation q;  LATENT_BWD: ml = [mu | logvar]
    const uint8_t* mask; int64_t ldmask; float keep_scale;
    const float* eps; int L; const float* coef;      // LATENT_BWD: eps [M, L], coef -> beta / B_global
    float* colsum; float* colsum_part; unsigned* colsum_ctr;    // column sums of what was written (nullable), partials [m_tiles][cols]
};
enum { TG_EPI_NONE = 0, TG_EPI_GELU_DROP = 1, TG_EPI_BF16 = 2, TG_EPI_GELU_BWD = 3, TG_EPI_LATENT_BWD = 4 };
// kernel instantiations: one epilogue fixed at compile time (>= 0), TG_EPI_PLAIN = NONE or BF16 chosen at run time (their code is the same
// size, so every plain GEMM of a step and the one that also writes the bf16 user vectors share ONE kernel), TG_EPI_ANY = all of them
enum { TG_EPI_ANY = -1, TG_EPI_PLAIN = -2 };
#define EPI_IS(X) ((EPI_T == (X) || EPI_T == TG_EPI_ANY || (EPI_T == TG_EPI_PLAIN && ((X) == TG_EPI_NONE || (X) == TG_EPI_BF16))) && EPI == (X))

struct __align__(8) TGemmBarriers {
    uint64_t raw_full[TG_MAX_STAGES], full[TG_MAX_STAGES], empty[TG_MAX_STAGES], acc_full;
    uint32_t tmem_base;
};

// raw [32 k][32 x] fp32 boxes (x = m or n, one box per 32 x) -> K-major tile rows x: 32 k-floats = 128 B, 16-byte chunk c
// stored at c ^ (x & 7) (the 128B swizzle the UMMA descriptor expects).  One thread per row x.
// Values are rounded to TF32 with round-to-nearest on the way (the MMA itself truncates fp32 operands, which biases
// every product low by ~2^-11: measured 1.5e-3 relative on the KL term).
__device__ __forceinline__ float rn_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// (All loads of a row are issued before its first store: chunk positions depend on the row index at run time, so the compiler
// cannot prove that a store does not alias a later load and would otherwise serialise 8 load -> store round trips per row.)
__device__ __forceinline__ void transpose_to_kmajor(const uint8_t* raw, uint8_t* tile, int x) {
    const float* src = reinterpret_cast<const float*>(raw + (x >> 5) * 4096) + (x & 31);
    uint8_t* dst = tile + x * 128;
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = src[k * 32];
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<float4*>(dst + ((c ^ (x & 7)) << 4)) =
            make_float4(rn_tf32(v[4 * c]), rn_tf32(v[4 * c + 1]), rn_tf32(v[4 * c + 2]), rn_tf32(v[4 * c + 3]));
}
// K-major operand already in place (TMA wrote it swizzled): round its row x in place.
__device__ __forceinline__ void round_row_inplace(uint8_t* tile, int x) {
    float4* row = reinterpret_cast<float4*>(tile + x * 128);
    float4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = row[c ^ (x & 7)];     // rows are 128 B apart: chunks in swizzled order -> bank-conflict free
#pragma unroll
    for (int c = 0; c < 8; ++c)
        row[c ^ (x & 7)] = make_float4(rn_tf32(v[c].x), rn_tf32(v[c].y), rn_tf32(v[c].z), rn_tf32(v[c].w));
}

// kind::tf32 instruction descriptor: D = f32, A = B = tf32 (format 2), major bits, N>>3, M>>4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) | (uint32_t(N >> 3) << 17) |
           (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// erf for the fused epilogues: erf(|x|) = 1 - 2^(-|x| P(|x|)), P a degree-7 minimax fit of -log2(erfc(t)) / t on [0, 4] (weighted for the
// absolute error of erf).  |error| < 1e-7 absolute over the whole line (the float32 rounding floor next to 1 is 6e-8), and no branch:
// the 32 independent evaluations of a row chunk interleave, which erff (two code paths) does not allow.  The epilogue runs on one or
// two warps per scheduler, so the dependent-issue latency of a branchy erff -- ~95 instructions per GELU -- was the whole cost of a
// fused epilogue (measured: +9 us on a 11 us GEMM at the C2 shape).
__device__ __forceinline__ float erf_fast(float x) {
    const float t = fminf(fabsf(x), 6.0f);
    float p = 4.5354507165029645e-05f;
    p = fmaf(p, t, -0.0004454739682842046f);
    p = fmaf(p, t, 0.0014893329935148358f);
    p = fmaf(p, t, 0.0007748050848022103f);
    p = fmaf(p, t, -0.02825383096933365f);
    p = fmaf(p, t, 0.14848168194293976f);
    p = fmaf(p, t, 0.9184163808822632f);
    p = fmaf(p, t, 1.6279085874557495f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-p * t));
    return copysignf(1.0f - e, x);
}
__device__ __forceinline__ float gelu_fast(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f)); }
// d/dx [x Phi(x)] = Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_grad_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));       // exp(-x^2 / 2)
    return fmaf(x * 0.39894228040143267794f, e, 0.5f * (1.0f + erf_fast(x * 0.70710678118654752440f)));
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }     // the eight epilogue warps

// Epilogue data layout.  The accumulator comes out of TMEM one ROW per thread; stores (and the loads of the fused modes) issued that
// way touch 32 different 128-byte lines per instruction, 16 bytes each -- the LSU spends 32 cycles on each of them and that, not the
// arithmetic, was what a fused epilogue cost (measured +17 us on a 20 us GEMM at 4096 x 768).  So every 32 x 32 chunk goes through a
// per-warp shared-memory tile [32][36] and is handled transposed: lane -> (row lane/8 + 4*it, columns 4*(lane%8)..+3), it = 0..7,
// i.e. each instruction covers four complete 128-byte row segments.
constexpr int TG_SCR_LD = 36, TG_SCR_BYTES = 32 * TG_SCR_LD * 4;

// four consecutive floats of a row at column `col` (zero beyond n_valid); 16-byte access when the row allows it
__device__ __forceinline__ void ld4(const float* p, int col, int n_valid, bool vec, float (&o)[4]) {
    if (vec && col + 3 < n_valid) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = col + e < n_valid ? p[e] : 0.f;
    }
}
__device__ __forceinline__ void st4(float* p, int col, int n_store, bool vec, const float (&o)[4]) {
    if (vec && col + 3 < n_store) {
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (col + e < n_store) p[e] = o[e];
    }
}
// four dropout keep bytes as one word
__device__ __forceinline__ uint32_t ldmask4(const uint8_t* p, int col, int n_valid) {
    if ((reinterpret_cast<uintptr_t>(p) & 3) == 0 && col + 3 < n_valid) return *reinterpret_cast<const uint32_t*>(p);
    uint32_t w = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) w |= (col + e < n_valid ? (uint32_t)p[e] : 0u) << (8 * e);
    return w;
}
// what a fused epilogue reads besides the accumulator for one 32 x 32 chunk (transposed layout); requested a chunk ahead: every global
// round trip on the epilogue's critical path costs ~0.5 us, a third of the whole main loop at the small shapes
struct EpiAux {
    float a[8][4];       // GELU_BWD: pre-activation q
    uint32_t mk[8];      // dropout mask bytes
};

// EPI_T >= 0: the epilogue is fixed at compile time; EPI_T < 0: one kernel serves every epilogue (P.epi).  These GEMMs run 8-15 us and
// execute most of their code once per CTA, so the SIZE of the kernel is part of its cost: the same plain GEMM takes ~2.4 us longer through
// the 145 KB all-epilogue kernel than through its own 58 KB instantiation (measured in the captured C2 step).
template <int TG_BN, int EPI_T>
__global__ void __launch_bounds__(TG_THREADS, 1) gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, TGemmParams P) {
    constexpr int TG_B_BYTES = TG_BN * TG_BK * 4, TG_TILES = TG_A_BYTES + TG_B_BYTES;
    const int EPI = EPI_T < 0 ? P.epi : EPI_T;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int TG_STAGES = P.n_stages, TG_STAGE = P.stage_bytes;
    TGemmBarriers* bars = reinterpret_cast<TGemmBarriers*>(smem + TG_SMEM_BUDGET);
    const int raw_a = TG_TILES, raw_b = TG_TILES + (P.a_mn ? TG_A_BYTES : 0);      // offsets of the raw areas in a stage
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * TG_BM, n0 = blockIdx.x * TG_BN;
    const int KB_all = (P.K + TG_BK - 1) / TG_BK;
    // split-K over a cluster of P.ksplit CTAs (these GEMMs are latency-bound per CTA: ~0.4 us per k-block, so the k loop is
    // shared out); partial accumulators are reduced through the leader's shared memory (DSMEM) in rank order: deterministic
    const int rank = P.ksplit > 1 ? (int)cluster_ctarank() : 0;
    const int kb_lo = (int)((int64_t)rank * KB_all / P.ksplit), kb_hi = (int)((int64_t)(rank + 1) * KB_all / P.ksplit);
    const int KB = kb_hi - kb_lo;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < TG_STAGES; ++s) { mbar_init(&bars->raw_full[s], 1); mbar_init(&bars->full[s], P.conv_per_stage); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TG_BN>(&bars->tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    // "dependents may launch" only now that this CTA holds its TMEM columns: a dependent tensor-core kernel that landed on this
    // SM earlier could take the columns and then park in griddepcontrol.wait on us while we block in tcgen05.alloc
    pdl_launch_dependents();
    pdl_wait();      // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel
    float* csum = reinterpret_cast<float*>(smem + TG_SMEM_BUDGET + 512);       // column-sum scratch [set][warp quarter][TG_BN]
    float* bias_s = csum + 8 * 256;                                            // this tile's bias values
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < TG_BN; i += 256) bias_s[i] = (P.bias && n0 + i < P.N) ? P.bias[n0 + i] : 0.f;
        epi_bar();
    }

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % TG_STAGES;
                mbar_wait(&bars->empty[s], ((kb / TG_STAGES) & 1) ^ 1);
                mbar_expect_tx(&bars->raw_full[s], TG_TILES);
                uint8_t* a = smem + s * TG_STAGE;
                uint8_t* b = a + TG_A_BYTES;
                const int kg = (kb_lo + kb) * TG_BK;                                      // global k of this block
                if (!P.a_mn) {
                    tma_load_2d(a, &tmA, kg, m0, &bars->raw_full[s]);                     // box {32 k, 128 m}, swizzled, final place
                } else {
                    for (int g = 0; g < TG_BM / 32; ++g)                                  // raw boxes {32 m, 32 k}
                        tma_load_2d(a + raw_a + g * 4096, &tmA, m0 + g * 32, kg, &bars->raw_full[s]);
                }
                if (!P.b_mn) {
                    tma_load_2d(b, &tmB, kg, n0, &bars->raw_full[s]);                     // box {32 k, 64 n}
                } else {
                    for (int g = 0; g < TG_BN / 32; ++g)                                  // raw boxes {32 n, 32 k}
                        tma_load_2d(a + raw_b + g * 4096, &tmB, n0 + g * 32, kg, &bars->raw_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(TG_BM, TG_BN, 0, 0);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % TG_STAGES;
                mbar_wait(P.direct ? &bars->raw_full[s] : &bars->full[s], (kb / TG_STAGES) & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(smem + s * TG_STAGE), b0 = a0 + TG_A_BYTES;
#pragma unroll
                for (int k = 0; k < TG_BK / 8; ++k)        // one MMA = 8 tf32 (32 B) along K; rows of 128 B, 8-row groups 1024 B apart
                    umma_tf32(tmem_base, make_desc(a0 + k * 32, 16, 1024), make_desc(b0 + k * 32, 16, 1024), idesc, (kb | k) != 0);
                umma_commit(&bars->empty[s]);
            }
            umma_commit(&bars->acc_full);
        }
    } else {
        // Main loop: the eight conversion warps each own every 8th k-block (their fence + barrier round trips overlap):
        // round (and, for MN-major operands, transpose) the whole stage into K-major TF32 tiles, lane <-> rows lane+32i.
        // (as many conversion warps as stages take part, so a stage is always served by the same warp: a parity wait
        // on a barrier that is a whole phase behind would otherwise pass immediately)
        // stage s is always served by the same g = conv_per_stage warps (a parity wait on a barrier that is a whole phase behind
        // would pass immediately); they share the stage's 32-row groups (A: BM/32, B: BN/32) round-robin
        const int cw = warp - 2, g = P.conv_per_stage;
        const int my_stage = cw / g, part = cw % g;
        for (int kb = my_stage; kb < KB && my_stage < TG_STAGES && !P.direct; kb += TG_STAGES) {
            const int s = kb % TG_STAGES;
            mbar_wait(&bars->raw_full[s], (kb / TG_STAGES) & 1);
            uint8_t* a = smem + s * TG_STAGE;
            // (rolled on purpose: one copy of each conversion routine instead of six -- these kernels execute most of their code once, so
            // code size is launch time)
#pragma unroll 1
            for (int i = part; i < TG_BM / 32 + TG_BN / 32; i += g) {
                const bool is_a = i < TG_BM / 32;
                const int x = lane + 32 * (is_a ? i : i - TG_BM / 32);
                uint8_t* tile = is_a ? a : a + TG_A_BYTES;
                if (is_a ? P.a_mn : P.b_mn) transpose_to_kmajor(a + (is_a ? raw_a : raw_b), tile, x);
                else round_row_inplace(tile, x);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->full[s]);
        }
    }
    // ---- epilogue ---------------------------------------------------------------------------------------------------
    // the eight conversion warps, two per TMEM lane quarter (a warp reaches lanes 32 * (warp % 4) ..), alternating 32-column chunks
    constexpr int NCH = TG_BN / 64;                // chunks per warp
    const bool gelu_mode = EPI_IS(TG_EPI_GELU_DROP) || EPI_IS(TG_EPI_GELU_BWD);
    const bool epi = warp >= 2;
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r_local = q * 32 + lane;             // accumulator row of this thread (TMEM lane)
    const int tr = lane >> 3, tcol = (lane & 7) * 4;       // transposed domain: rows tr + 4*it, columns tcol..tcol+3 of the chunk
    const int mq = m0 + q * 32;                    // first row of this warp's quarter
    const bool lead = epi && rank == 0;
    const bool has_mask = P.mask != nullptr;
    const bool aux_vec = P.ldaux % 4 == 0 && (reinterpret_cast<uintptr_t>(P.aux) & 15) == 0;
    auto load_aux = [&](EpiAux& x, int c) {
        const int col = n0 + c * 32 + tcol;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int mg = mq + it * 4 + tr;
            if (mg >= P.M) continue;
            if (EPI_IS(TG_EPI_GELU_BWD)) ld4(P.aux + (int64_t)mg * P.ldaux + col, col, P.N, aux_vec, x.a[it]);
            if (has_mask) x.mk[it] = ldmask4(P.mask + (int64_t)mg * P.ldmask + col, col, P.N);
        }
    };
    EpiAux cur;
    if (lead && gelu_mode) load_aux(cur, half);    // operands of the first chunk are requested before the accumulator is complete
    if (epi) {
        mbar_wait(&bars->acc_full, 0);
        tc_fence_after();
    }
    if (P.ksplit > 1) {
        cluster_sync_all();        // every CTA of the cluster has finished its k range: the leader's ring buffers are free
        if (epi && rank > 0) {     // partial accumulator -> leader's shared memory, [rank-1][16-byte column chunk][row]
            const uint32_t red = map_to_cta(smem_u32(smem), 0u) + (uint32_t)(rank - 1) * (TG_BN / 4) * TG_BM * 16;
#pragma unroll 1
            for (int c = half; c < TG_BN / 32; c += 2) {
                float v[32];
                tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    st_cluster_v4(red + (uint32_t)(((c * 8 + j / 4) * TG_BM + r_local) * 16), __float_as_uint(v[j]), __float_as_uint(v[j + 1]),
                                  __float_as_uint(v[j + 2]), __float_as_uint(v[j + 3]));
            }
        }
        cluster_sync_all();        // partials are visible in the leader
    }
    if (lead) {
        const bool vec = (P.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
        const float4* red = reinterpret_cast<const float4*>(smem);
        // transposition tile of this warp: the tail of the operand ring (free: every stage was consumed; the split-K partials above
        // start at the head of the ring and the host keeps them clear of the tail)
        float* scr = reinterpret_cast<float*>(smem + TG_SMEM_BUDGET - 8 * TG_SCR_BYTES) + (warp - 2) * (TG_SCR_BYTES / 4);
        const bool want_cs = P.colsum != nullptr;
        const int n_sets = EPI_IS(TG_EPI_LATENT_BWD) ? 2 : 1;
        const float ks = has_mask ? P.keep_scale : 1.f;
#pragma unroll 1
        for (int ci = 0; ci < NCH; ++ci) {
            const int c = half + 2 * ci;
            const int col = n0 + c * 32 + tcol;          // first of this lane's four columns
            if (ci > 0 && gelu_mode) load_aux(cur, c);
            float cs[4], cs2[4];                         // column sums of this lane's rows
            // LATENT_BWD: logvar and eps rows are requested here, before the accumulator chunk is waited for; mu follows once their
            // registers are free
            float lv[8][4], ep[8][4];
            if (EPI_IS(TG_EPI_LATENT_BWD)) {
                const int L = P.L;
                const bool evec = L % 4 == 0 && (reinterpret_cast<uintptr_t>(P.eps) & 15) == 0;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int mg = mq + it * 4 + tr;
                    if (mg >= P.M) continue;
                    ld4(P.aux + (int64_t)mg * P.ldaux + L + col, col, L, aux_vec && L % 4 == 0, lv[it]);
                    if (P.eps) ld4(P.eps + (int64_t)mg * L + col, col, L, evec, ep[it]);
                }
            }
            {
                float v[32];
                tmem_ld32(tmem_base + (uint32_t(q * 32) << 16) + c * 32, v);
                tmem_ld_wait();
                for (int pr = 0; pr < P.ksplit - 1; ++pr) {          // fixed rank order
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 t = red[(size_t)pr * (TG_BN / 4) * TG_BM + (c * 8 + j / 4) * TG_BM + r_local];
                        v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(scr + lane * TG_SCR_LD + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            __syncwarp();
            float b4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { b4[e] = bias_s[c * 32 + tcol + e]; cs[e] = cs2[e] = 0.f; }
            // (the modes without prefetched operands keep this loop rolled: smaller code, see above)
            constexpr int IT_UNROLL = (EPI_T == TG_EPI_NONE || EPI_T == TG_EPI_BF16 || EPI_T == TG_EPI_PLAIN) ? 1 : 8;
#pragma unroll IT_UNROLL
            for (int it = 0; it < 8; ++it) {
                const int rr = it * 4 + tr, mg = mq + rr;
                if (mg >= P.M) continue;
                float x[4];
                {
                    const float4 t = *reinterpret_cast<const float4*>(scr + rr * TG_SCR_LD + tcol);
                    x[0] = fmaf(t.x, P.alpha, b4[0]); x[1] = fmaf(t.y, P.alpha, b4[1]); x[2] = fmaf(t.z, P.alpha, b4[2]); x[3] = fmaf(t.w, P.alpha, b4[3]);
                }
                float* crow = P.C + (int64_t)mg * P.ldc + col;
                if (EPI_IS(TG_EPI_NONE)) {
                    st4(crow, col, P.n_store, vec, x);
                } else if (EPI_IS(TG_EPI_GELU_DROP)) {
                    st4(crow, col, P.n_store, vec, x);
                    const uint32_t mk = has_mask ? cur.mk[it] : 0x01010101u;
#pragma unroll
                    for (int e = 0; e < 4; ++e) x[e] = (col + e < P.N && ((mk >> (8 * e)) & 0xffu)) ? gelu_fast(x[e]) * ks : 0.f;
                    st4(P.C2 + (int64_t)mg * P.ldc + col, col, P.n_store, vec, x);
                } else if (EPI_IS(TG_EPI_BF16)) {
                    if (P.C) st4(crow, col, P.n_store, vec, x);
                    __nv_bfloat16* brow = P.Cb + (int64_t)mg * P.ldcb + col;
                    if (P.ldcb % 4 == 0 && (reinterpret_cast<uintptr_t>(P.Cb) & 7) == 0 && col + 3 < P.ldcb) {
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(x[0], x[1]), hi = __floats2bfloat162_rn(x[2], x[3]);
                        *reinterpret_cast<uint2*>(brow) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (col + e < P.ldcb) brow[e] = __float2bfloat16(x[e]);
                    }
                } else if (EPI_IS(TG_EPI_GELU_BWD)) {
                    const uint32_t mk = has_mask ? cur.mk[it] : 0x01010101u;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        x[e] = (col + e < P.N && ((mk >> (8 * e)) & 0xffu)) ? x[e] * ks * gelu_grad_fast(cur.a[it][e]) : 0.f;
                        cs[e] += x[e];
                    }
                    st4(crow, col, P.n_store, vec, x);
                } else if (EPI_IS(TG_EPI_LATENT_BWD)) {      // first half: x = dz -> d logvar (model.py:157-179 backward + KL gradient)
                    const int L = P.L;
                    const float cf = *P.coef;
                    float g[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float sdev = expf(0.5f * lv[it][e]);                    // std; exp(logvar) = std^2
                        float t = cf * 0.5f * (sdev * sdev - 1.0f);
                        if (P.eps) t += x[e] * ep[it][e] * 0.5f * sdev;
                        g[e] = col + e < L ? t : 0.f;
                        cs2[e] += g[e];
                    }
                    st4(crow + L, col, L, vec && L % 4 == 0, g);
                }
            }
            if (EPI_IS(TG_EPI_LATENT_BWD)) {      // second half: d mu = dz + coef * mu
                const int L = P.L;
                const float cf = *P.coef;
                float mu[8][4];
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int mg = mq + it * 4 + tr;
                    if (mg < P.M) ld4(P.aux + (int64_t)mg * P.ldaux + col, col, L, aux_vec && L % 4 == 0, mu[it]);
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + tr, mg = mq + rr;
                    if (mg >= P.M) continue;
                    const float4 t = *reinterpret_cast<const float4*>(scr + rr * TG_SCR_LD + tcol);
                    float x[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        x[e] = col + e < L ? x[e] + cf * mu[it][e] : 0.f;
                        cs[e] += x[e];
                    }
                    st4(P.C + (int64_t)mg * P.ldc + col, col, L, vec && L % 4 == 0, x);
                }
            }
            __syncwarp();          // the tile is rewritten by the next chunk
            if (want_cs) {         // lanes with the same columns (lane % 8) add up their rows: the warp's 32 column sums of this chunk
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float t = cs[e];
                    t += __shfl_xor_sync(0xffffffffu, t, 8);
                    t += __shfl_xor_sync(0xffffffffu, t, 16);
                    if (tr == 0) csum[q * TG_BN + c * 32 + tcol + e] = t;
                    if (n_sets == 2) {
                        float u = cs2[e];
                        u += __shfl_xor_sync(0xffffffffu, u, 8);
                        u += __shfl_xor_sync(0xffffffffu, u, 16);
                        if (tr == 0) csum[(4 + q) * TG_BN + c * 32 + tcol + e] = u;
                    }
                }
            }
        }
        if (want_cs) {
            // bias gradient: the tile's column sums go to a partial row of this m tile; the last m tile to arrive adds the partial rows
            // up in tile order
            const int e = ((warp - 2) << 5) + lane, m_tiles = (int)gridDim.y;       // 0..255
            const int cols_all = EPI_IS(TG_EPI_LATENT_BWD) ? 2 * P.L : P.N, n_lim = EPI_IS(TG_EPI_LATENT_BWD) ? P.L : P.N;
            __shared__ int s_last;
            epi_bar();
            for (int i = e; i < n_sets * TG_BN; i += 256) {
                const int set = i / TG_BN, colx = i - set * TG_BN;
                const float* cp = csum + set * 4 * TG_BN + colx;
                const float t = (cp[0] + cp[TG_BN]) + (cp[2 * TG_BN] + cp[3 * TG_BN]);          // the four 32-row quarters of the tile
                if (n0 + colx < n_lim) __stcg(&P.colsum_part[(size_t)blockIdx.y * cols_all + set * P.L + n0 + colx], t);
            }
            __threadfence();
            epi_bar();
            if (e == 0) s_last = atomicAdd(&P.colsum_ctr[blockIdx.x], 1u) == (unsigned)(m_tiles - 1);
            epi_bar();
            if (s_last) {
                __threadfence();
                for (int i = e; i < n_sets * TG_BN; i += 256) {
                    const int set = i / TG_BN, colx = i - set * TG_BN;
                    if (n0 + colx >= n_lim) continue;
                    const size_t g = (size_t)set * P.L + n0 + colx;
                    float t = 0.f;
                    for (int mt0 = 0; mt0 < m_tiles; mt0 += 8) {            // eight loads in flight, added in tile order
                        float x[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) x[u] = mt0 + u < m_tiles ? __ldcg(P.colsum_part + (size_t)(mt0 + u) * cols_all + g) : 0.f;
#pragma unroll
                        for (int u = 0; u < 8; ++u) t += x[u];
                    }
                    P.colsum[g] = t;
                }
                if (e == 0) P.colsum_ctr[blockIdx.x] = 0u;       // ready for the next launch
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TG_BN>(tmem_base);
    }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// fp32 2-D view [outer, inner] (inner contiguous, outer stride `ld` floats); box = {32 inner, box_outer}; 128B swizzle
static int make_tmap_f32(CUtensorMap* out, const float* base, int64_t inner, int64_t outer, int64_t ld, int box_outer, bool swizzle) {
    static EncodeTiledFn2 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn2>(p);
        if (!fn) return hvae_fail("cuTensorMapEncodeTiled is not available from the driver");
    }
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return hvae_fail("cuTensorMapEncodeTiled(f32) failed (%d) inner=%lld outer=%lld ld=%lld", (int)r, (long long)inner,
                                            (long long)outer, (long long)ld);
    return 0;
}

constexpr size_t kTGemmSmem = TG_SMEM_BUDGET + 512 + 8 * 256 * 4 + 256 * 4 + 1024;     // ring | barriers | column-sum scratch | bias | alignment slack

}  // namespace tc
}  // namespace hvae

using namespace hvae;
using namespace hvae::tc;

// One launch configuration of the templated kernel.
template <int BN, int EPI>
static int launch_gemm_tf32_epi(TGemmParams P, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                            int ksplit, cudaStream_t stream) {
    constexpr int B_BYTES = BN * TG_BK * 4, TILES = TG_A_BYTES + B_BYTES;
    const int M = P.M, N = P.N, K = P.K;
    P.stage_bytes = TILES + (P.a_mn ? TG_A_BYTES : 0) + (P.b_mn ? B_BYTES : 0);
    P.n_stages = min(TG_MAX_STAGES, TG_SMEM_BUDGET / P.stage_bytes);
    P.conv_per_stage = max(1, TG_CONV_WARPS / P.n_stages);
    P.ksplit = ksplit;
    CUtensorMap tmA, tmB;
    // A(m,k): K-major -> view [M outer, K inner] stride a_rs; MN-major -> view [K outer, M inner] stride a_cs
    if (int rc = P.a_mn ? make_tmap_f32(&tmA, A, M, K, a_cs, 32, false) : make_tmap_f32(&tmA, A, K, M, a_rs, TG_BM, true)) return rc;
    // B(k,n): K-major (k contiguous) -> view [N outer, K inner] stride b_cs; MN-major -> view [K outer, N inner] stride b_rs
    if (int rc = P.b_mn ? make_tmap_f32(&tmB, B, N, K, b_rs, 32, false) : make_tmap_f32(&tmB, B, K, N, b_cs, BN, true)) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        HVAE_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTGemmSmem));
        attr_set = true;
    }
    dim3 grid(ceil_div(N, BN), ceil_div(M, TG_BM), ksplit);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(TG_THREADS);
    cfg.dynamicSmemBytes = kTGemmSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (ksplit > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = ksplit;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    HVAE_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<BN, EPI>, tmA, tmB, P));
    HVAE_LAUNCH_CHECK("gemm_tf32");
    return 0;
}

// One instantiation per epilogue (EPI_T >= 0): the plain GEMM keeps its small code.  HVAE_TF32_ONE_KERNEL=1 routes every epilogue through
// the EPI_T = -1 instantiation instead (profiles/r02_ncu_summary.md section 7 has both measured).
template <int BN>
static int launch_gemm_tf32(TGemmParams P, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                            int ksplit, cudaStream_t stream) {
    static const bool one_kernel = getenv("HVAE_TF32_ONE_KERNEL") != nullptr;
    if (one_kernel) return launch_gemm_tf32_epi<BN, TG_EPI_ANY>(P, A, a_rs, a_cs, B, b_rs, b_cs, ksplit, stream);
    switch (P.epi) {
        case TG_EPI_GELU_DROP: return launch_gemm_tf32_epi<BN, TG_EPI_GELU_DROP>(P, A, a_rs, a_cs, B, b_rs, b_cs, ksplit, stream);
        case TG_EPI_GELU_BWD: return launch_gemm_tf32_epi<BN, TG_EPI_GELU_BWD>(P, A, a_rs, a_cs, B, b_rs, b_cs, ksplit, stream);
        case TG_EPI_LATENT_BWD: return launch_gemm_tf32_epi<BN, TG_EPI_LATENT_BWD>(P, A, a_rs, a_cs, B, b_rs, b_cs, ksplit, stream);
        default: return launch_gemm_tf32_epi<BN, TG_EPI_PLAIN>(P, A, a_rs, a_cs, B, b_rs, b_cs, ksplit, stream);       // NONE, BF16
    }
}

static size_t tf32_operands_ok(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs) {
    const bool a_ok = (a_cs == 1 && a_rs % 4 == 0) || (a_rs == 1 && a_cs % 4 == 0);
    const bool b_ok = (b_rs == 1 && b_cs % 4 == 0) || (b_cs == 1 && b_rs % 4 == 0);
    return (a_ok && b_ok && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0) ? 1 : 0;
}

// Tile width and split-K choice, then the launch.  P carries the outputs and the epilogue; the operand description is filled here.
static int gemm_tf32_dispatch(TGemmParams P, int M, int N, int K, const float* A, int64_t a_rs, int64_t a_cs, const float* B,
                              int64_t b_rs, int64_t b_cs, cudaStream_t stream) {
    if (M == 0 || N == 0) return 0;
    HVAE_REQUIRE(K > 0, "gemm_tf32: K must be positive");
    HVAE_REQUIRE(tf32_operands_ok(A, a_rs, a_cs, B, b_rs, b_cs), "gemm_tf32: operand strides/alignment not TMA-compatible");
    P.M = M; P.N = N; P.K = K;
    P.a_mn = (a_cs == 1) ? 0 : 1;
    P.b_mn = (b_rs == 1) ? 0 : 1;
    if (a_cs == 1 && a_rs == 1) P.a_mn = 0;
    if (b_rs == 1 && b_cs == 1) P.b_mn = 0;
    static const bool trunc = getenv("HVAE_TF32_TRUNC") != nullptr;
    P.direct = (trunc && !P.a_mn && !P.b_mn) ? 1 : 0;
    static const bool no_split = getenv("HVAE_NO_SPLITK") != nullptr;
    static const char* force_bn = getenv("HVAE_TF32_BN");
    const int KB = ceil_div(K, TG_BK), m_tiles = ceil_div(M, TG_BM);
    // Tile width (measured on B200 at the C2 / C3 shapes, tools/bench_gemm_c3.py): 128 x 64 tiles for small batches (more CTAs,
    // split-K over a cluster); from 2,048 rows (or k's) on, 128 x 256 when both operands are K-major and N >= 512 (4-stage ring: the
    // rounding pass over the A tile is shared by 4x the MMA work), 128 x 128 when there are >= 24 such tiles and N >= 384 (an
    // MN-major operand needs a raw staging area per stage: at 256 columns only 2 stages fit), else 128 x 64.
    const int big = max(M, K) >= 2048;
    int best_bn = 64;
    if (big && !P.a_mn && !P.b_mn && N >= 512) best_bn = 256;
    else if (big && N >= 384 && m_tiles * ceil_div(N, 128) >= 24) best_bn = 128;
    if (force_bn) best_bn = atoi(force_bn);
    HVAE_REQUIRE(best_bn == 64 || best_bn == 128 || best_bn == 256, "gemm_tf32: HVAE_TF32_BN must be 64, 128 or 256");
    // Split-K over a cluster: a CTA costs ~5.6 us fixed + ~0.4-0.65 us per 32-wide k-block, a split adds ~1 us (two cluster barriers
    // + the DSMEM reduction); only while every CTA keeps >= 4 k-blocks, the partials fit shared memory and the grid stays in one wave.
    const int tiles = m_tiles * ceil_div(N, best_bn);
    const int max_split = no_split ? 1 : min(4, (TG_SMEM_BUDGET - 8 * TG_SCR_BYTES) / (best_bn * TG_BM * 4) + 1);     // partials stay clear of the epilogue's transposition tiles
    int best_split = 1;
    if (max_split >= 4 && KB >= 16 && tiles * 4 <= 132) best_split = 4;
    else if (max_split >= 2 && KB >= 8 && tiles * 2 <= 132) best_split = 2;
    if (best_bn == 256) return launch_gemm_tf32<256>(P, A, a_rs, a_cs, B, b_rs, b_cs, best_split, stream);
    if (best_bn == 128) return launch_gemm_tf32<128>(P, A, a_rs, a_cs, B, b_rs, b_cs, best_split, stream);
    return launch_gemm_tf32<64>(P, A, a_rs, a_cs, B, b_rs, b_cs, best_split, stream);
}

extern "C" {

// 1 if the operands satisfy the TMA constraints of hvae_gemm_tf32 (unit stride on one axis of A and of B, the
// other stride a multiple of 4 floats, 16-byte aligned bases), else 0.
size_t hvae_gemm_tf32_supported(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs) {
    return tf32_operands_ok(A, a_rs, a_cs, B, b_rs, b_cs);
}

int hvae_gemm_tf32(int M, int N, int K, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                   float* C, int64_t ldc, const float* bias, float alpha, void* stream) {
    TGemmParams P{};
    P.C = C; P.ldc = ldc; P.bias = bias; P.alpha = alpha; P.epi = TG_EPI_NONE; P.n_store = N;
    return gemm_tf32_dispatch(P, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, (cudaStream_t)stream);
}

// ---- the same GEMM with the element-wise kernel that follows it in the model folded into the epilogue --------------------------

size_t hvae_gemm_colsum_workspace_floats(int M, int cols) { return (size_t)ceil_div(M, TG_BM) * cols + ((ceil_div(cols, 64) + 63) / 64) * 64; }

// column-sum workspace: [counters: one per n tile, zero on first use, the kernel leaves them zero][partials m_tiles x cols]
static void set_colsum(TGemmParams& P, float* colsum, float* workspace, int cols) {
    P.colsum = colsum;
    P.colsum_ctr = reinterpret_cast<unsigned*>(workspace);
    P.colsum_part = workspace + ((ceil_div(cols, 64) + 63) / 64) * 64;
}

int hvae_gemm_tf32_gelu_drop(int M, int N, int K, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                             float* pre, float* act, int64_t ldc, const float* bias, const uint8_t* mask, float keep_scale,
                             void* stream) {
    HVAE_REQUIRE(ldc >= N && (reinterpret_cast<uintptr_t>(pre) & 15) == (reinterpret_cast<uintptr_t>(act) & 15), "gemm_tf32_gelu_drop: bad outputs");
    TGemmParams P{};
    P.C = pre; P.C2 = act; P.ldc = ldc; P.bias = bias; P.alpha = 1.f; P.epi = TG_EPI_GELU_DROP; P.n_store = (int)ldc;
    P.mask = mask; P.ldmask = N; P.keep_scale = keep_scale;
    return gemm_tf32_dispatch(P, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, (cudaStream_t)stream);
}

int hvae_gemm_tf32_bf16(int M, int N, int K, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                        float* C, int64_t ldc, const float* bias, void* C_bf16, int64_t ld_bf16, void* stream) {
    HVAE_REQUIRE(C_bf16 && ld_bf16 >= N && (!C || ldc >= N), "gemm_tf32_bf16: bad outputs");
    TGemmParams P{};
    P.C = C; P.ldc = ldc; P.bias = bias; P.alpha = 1.f; P.epi = TG_EPI_BF16; P.n_store = (int)ldc;
    P.Cb = (__nv_bfloat16*)C_bf16; P.ldcb = ld_bf16;
    HVAE_REQUIRE(ld_bf16 <= (int64_t)ceil_div(N, 64) * 64, "gemm_tf32_bf16: bf16 padding beyond the last tile");
    return gemm_tf32_dispatch(P, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, (cudaStream_t)stream);
}

int hvae_gemm_tf32_gelu_bwd(int M, int N, int K, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                            float* dpre, int64_t ldc, const float* pre, const uint8_t* mask, float keep_scale, float* colsum,
                            float* workspace, void* stream) {
    HVAE_REQUIRE(ldc >= N && (!colsum || workspace), "gemm_tf32_gelu_bwd: bad arguments");
    TGemmParams P{};
    P.C = dpre; P.ldc = ldc; P.alpha = 1.f; P.epi = TG_EPI_GELU_BWD; P.n_store = (int)ldc;
    P.aux = pre; P.ldaux = ldc; P.mask = mask; P.ldmask = N; P.keep_scale = keep_scale;
    if (colsum) set_colsum(P, colsum, workspace, N);
    return gemm_tf32_dispatch(P, M, N, K, A, a_rs, a_cs, B, b_rs, b_cs, (cudaStream_t)stream);
}

int hvae_gemm_tf32_latent_bwd(int M, int L, int K, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                              const float* ml, int64_t ldml, const float* eps, const float* coef, float* dml, float* colsum,
                              float* workspace, void* stream) {
    HVAE_REQUIRE(ldml >= 2 * L && coef && (!colsum || workspace), "gemm_tf32_latent_bwd: bad arguments");
    TGemmParams P{};
    P.C = dml; P.ldc = ldml; P.alpha = 1.f; P.epi = TG_EPI_LATENT_BWD; P.n_store = L;
    P.aux = ml; P.ldaux = ldml; P.eps = eps; P.L = L; P.coef = coef;
    if (colsum) set_colsum(P, colsum, workspace, 2 * L);
    return gemm_tf32_dispatch(P, M, L, K, A, a_rs, a_cs, B, b_rs, b_cs, (cudaStream_t)stream);
}

}  // extern "C"
