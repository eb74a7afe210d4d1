// Small dense GEMMs of the MLP stack on the tensor cores (bf16/"tensor-core" mode of the model):
//   C[m,n] = alpha * sum_k A(m,k) * B(k,n) (+ bias[n]),   fp32 operands read as TF32, fp32 accumulation in TMEM.
// Same contract as hvae_gemm_f32 (arbitrary transposes through strides), so the engine swaps one for the other:
// fc_mu/fc_logvar, the projection MLP and the deeper hidden layers, forward, dX and dW
// (reference src/ml/model.py:90-95,114,126-127 and their autograd).  At a 512-user batch these nine GEMMs are
// 400-600 wide and latency-bound; a TMA-fed tcgen05 tile (128 x 64, k-blocks of 32 floats = one 128B swizzle span)
// turns each into a few microseconds.  The fp32 ("exact") mode keeps the FFMA kernel.
// The MMA reads both operands K-major (128B-swizzled rows of 32 floats).  An operand whose m / n axis is the
// contiguous one in memory (the transposed uses: dX = dY W, dW = dY^T X) is TMA-loaded un-swizzled into a staging
// area and transposed smem -> smem by the four epilogue warps, which are idle during the main loop anyway
// (kind::tf32 does not accept MN-major shared-memory descriptors in the 128B-swizzle form: measured, it yields zeros).
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
};

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

template <int TG_BN>
__global__ void __launch_bounds__(TG_THREADS, 1) gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmB, TGemmParams P) {
    constexpr int TG_B_BYTES = TG_BN * TG_BK * 4, TG_TILES = TG_A_BYTES + TG_B_BYTES;
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
#pragma unroll
            for (int i = 0; i < TG_BM / 32 + TG_BN / 32; ++i) {
                if (i % g != part) continue;
                if (i < TG_BM / 32) {
                    const int x = lane + 32 * i;
                    if (P.a_mn) transpose_to_kmajor(a + raw_a, a, x);
                    else round_row_inplace(a, x);
                } else {
                    const int x = lane + 32 * (i - TG_BM / 32);
                    if (P.b_mn) transpose_to_kmajor(a + raw_b, a + TG_A_BYTES, x);
                    else round_row_inplace(a + TG_A_BYTES, x);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->full[s]);
        }
    }
    // ---- epilogue ---------------------------------------------------------------------------------------------------
    const bool epi = warp >= 2 && warp < 6;          // four warps, one per TMEM lane quarter
    const int q = warp & 3;
    const int r_local = q * 32 + lane;
    const int m = m0 + r_local;
    if (epi) {
        mbar_wait(&bars->acc_full, 0);
        tc_fence_after();
    }
    if (P.ksplit > 1) {
        cluster_sync_all();        // every CTA of the cluster has finished its k range: the leader's ring buffers are free
        if (epi && rank > 0) {     // partial accumulator -> leader's shared memory, [rank-1][16-byte column chunk][row]
            const uint32_t red = map_to_cta(smem_u32(smem), 0u) + (uint32_t)(rank - 1) * (TG_BN / 4) * TG_BM * 16;
#pragma unroll 1
            for (int c = 0; c < TG_BN / 32; ++c) {
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
    if (epi && rank == 0) {
        const bool vec = (P.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(P.C) & 15) == 0);
        const float4* red = reinterpret_cast<const float4*>(smem);
#pragma unroll 1
        for (int c = 0; c < TG_BN / 32; ++c) {
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
            if (m >= P.M) continue;
            const int nb = n0 + c * 32;
            float* crow = P.C + (int64_t)m * P.ldc + nb;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int n = nb + j + e;
                    o[e] = v[j + e] * P.alpha + ((P.bias && n < P.N) ? P.bias[n] : 0.f);
                }
                if (vec && nb + j + 3 < P.N) {
                    *reinterpret_cast<float4*>(crow + j) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (nb + j + e < P.N) crow[j + e] = o[e];
                }
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

constexpr size_t kTGemmSmem = TG_SMEM_BUDGET + 512 + 1024;

}  // namespace tc
}  // namespace hvae

using namespace hvae;
using namespace hvae::tc;

// One launch configuration of the templated kernel.
template <int BN>
static int launch_gemm_tf32(TGemmParams P, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
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
        HVAE_CUDA(cudaFuncSetAttribute(gemm_tf32_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTGemmSmem));
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
    HVAE_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<BN>, tmA, tmB, P));
    HVAE_LAUNCH_CHECK("gemm_tf32");
    return 0;
}

extern "C" {

// 1 if the operands satisfy the TMA constraints of hvae_gemm_tf32 (unit stride on one axis of A and of B, the
// other stride a multiple of 4 floats, 16-byte aligned bases), else 0.
size_t hvae_gemm_tf32_supported(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs) {
    const bool a_ok = (a_cs == 1 && a_rs % 4 == 0) || (a_rs == 1 && a_cs % 4 == 0);
    const bool b_ok = (b_rs == 1 && b_cs % 4 == 0) || (b_cs == 1 && b_rs % 4 == 0);
    return (a_ok && b_ok && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0) ? 1 : 0;
}

int hvae_gemm_tf32(int M, int N, int K, const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                   float* C, int64_t ldc, const float* bias, float alpha, void* stream) {
    if (M == 0 || N == 0) return 0;
    HVAE_REQUIRE(K > 0, "gemm_tf32: K must be positive");
    HVAE_REQUIRE(hvae_gemm_tf32_supported(A, a_rs, a_cs, B, b_rs, b_cs), "gemm_tf32: operand strides/alignment not TMA-compatible");
    TGemmParams P{};
    P.M = M; P.N = N; P.K = K; P.C = C; P.ldc = ldc; P.bias = bias; P.alpha = alpha;
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
    const int max_split = no_split ? 1 : min(4, TG_SMEM_BUDGET / (best_bn * TG_BM * 4) + 1);
    int best_split = 1;
    if (max_split >= 4 && KB >= 16 && tiles * 4 <= 132) best_split = 4;
    else if (max_split >= 2 && KB >= 8 && tiles * 2 <= 132) best_split = 2;
    if (best_bn == 256) return launch_gemm_tf32<256>(P, A, a_rs, a_cs, B, b_rs, b_cs, best_split, (cudaStream_t)stream);
    if (best_bn == 128) return launch_gemm_tf32<128>(P, A, a_rs, a_cs, B, b_rs, b_cs, best_split, (cudaStream_t)stream);
    return launch_gemm_tf32<64>(P, A, a_rs, a_cs, B, b_rs, b_cs, best_split, (cudaStream_t)stream);
}

}  // extern "C"
