// The two chained scoring GEMMs (S = U E_t^T  ->  P = exp(S - shift)  ->  O += P E_t, see score_tc.cu) on CTA PAIRS:
// tcgen05.mma.cta_group::2 with M = 128 users per pair, 64 per CTA.  Included by score_tc.cu (same namespace and parameters).
//
// Why: the one-CTA-per-column-chunk kernels above re-stream the user tile for every item tile and read every E tile once per
// GEMM and column chunk -- 576 KB L2->SM per CTA per two 128-item tiles at d = 768, 94 B per tensor cycle, and the M = N = 128
// MMAs read 128 B/cycle of shared memory on top of the TMA writes: tensor pipe 39 % active.  Here
//   * CTA r of the pair owns users [64 r, 64 r + 64) of the 128-user tile -- for ALL d columns: with M = 128 over two CTAs the
//     accumulator rows of a CTA sit in TMEM lanes 0..63 for the first N/2 columns and in lanes 64..127 for the second N/2
//     ("2x2" data-path layout), so O [64 x 768] takes 384 TMEM columns and an S tile [64 x 128] takes 64;
//   * its U rows (64 x d bf16 <= 96 KB) stay RESIDENT in shared memory for the whole sweep (A operand of G1);
//   * the B operands are split between the two CTAs by the hardware: for G1 each CTA loads 64 of the tile's 128 items (all d
//     columns), for G2 each loads all 128 items for half of the O columns -- 192 KB per CTA per 128-item tile instead of 288,
//     and the shared-memory reads per MMA halve (each SM reads its own half of B);
//   * P never leaves the CTA: each CTA's softmax threads produce exactly the A rows its own half of the G2 MMA needs
//     (the pair kernel above pushed every P tile through DSMEM).
// Roles per CTA: warp 0 = TMA producer (its own boxes; completion bytes credited to CTA 0's barriers), warp 1 = MMA issuer
// (CTA 0 only; both allocate TMEM), warps 2..5 = softmax / epilogue (thread <-> TMEM lane: user = lane & 63, item half = lane >> 6).
// Ring (7 stages) stage = two [64 x 64] bf16 boxes (16 KB) per CTA: two k-blocks of the CTA's 64 items for G1, two 64-column boxes of a
// 64-item half for G2.  Tensor-pipe order: G1(0), G1(1), G2(0), G1(2), G2(1), ... -- the softmax of tile t hides behind G1(t+1).
#pragma once

constexpr int D_BOX = 8192;                    // one [64 rows x 64 columns] bf16 box, 128B swizzle
constexpr int D_STAGE = 2 * D_BOX;
constexpr int D_STAGES = 7;
constexpr int D_MAXD = 768;                    // 12 resident U boxes; O = 3 groups x 128 TMEM columns
constexpr int D_UBYTES = (D_MAXD / 64) * D_BOX;
constexpr int D_PBYTES = 2 * D_BOX;            // P tile [64 users x 128 items] bf16 = two 64-item swizzle atoms
constexpr int D_TMEM_S = 384;                  // S buffers at TMEM columns 384 and 448

struct __align__(8) DuoBarriers {
    uint64_t full[D_STAGES], empty[D_STAGES], u_full, s_full[2], s_free[2], p_full[2], p_free[2], o_full;
    uint32_t tmem_base;
    uint32_t vote[2];         // sweep-repeat votes of the two CTAs (each CTA holds both)
    float rsum[2][64];        // row sums of the two item halves of this CTA's users
};

// Softmax numerators of one thread's 64 scores (one user, one 64-item half of the tile): p = exp2(v * log2e - shift2), rounded to
// bf16 and stored as the thread's 128-byte row of a 128B-swizzled [64 x 64] A-operand atom.  Returns the sum of the unrounded
// numerators over the first n_valid items (MASK: the catalogue ends inside this half).
// 2^x on the SFU, results below 2^-126 flushed to zero (exp2f() adds a compare and two predicated multiplies per element to
// produce denormals nobody needs here: the sweep-repeat logic already treats row sums below 2^-50 as underflow)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool MASK>
__device__ __forceinline__ float duo_softmax_row(const float (&v)[2][32], float shift2, int n_valid, uint8_t* prow, int u_local) {
    float ls = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {         // 8 items -> one 16-byte chunk
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int idx = c * 32 + j8 * 8 + 2 * e;
                float p0 = ex2_ftz(fmaf(v[c][j8 * 8 + 2 * e], kLog2e, -shift2));
                float p1 = ex2_ftz(fmaf(v[c][j8 * 8 + 2 * e + 1], kLog2e, -shift2));
                if (MASK) {
                    if (idx >= n_valid) p0 = 0.f;
                    if (idx + 1 >= n_valid) p1 = 0.f;
                }
                ls += p0 + p1;
                __nv_bfloat162 h2 = __floats2bfloat162_rn(p0, p1);
                w[e] = *reinterpret_cast<uint32_t*>(&h2);
            }
            const int chunk16 = c * 4 + j8;
            *reinterpret_cast<uint4*>(prow + ((chunk16 ^ (u_local & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    return ls;
}

constexpr size_t kDuoSmem = 1024 + D_UBYTES + D_STAGES * D_STAGE + D_PBYTES + 1024;

// TRACE (profiling builds of the same code, hvae_tc_duo_trace): lane 0 of the producer / MMA / first softmax warp accumulates
// clock64() cycles spent in each of its waits and writes 8 counters per role and CTA to `trace`.
#define DUO_TIMED(slot, stmt)                                  \
    do {                                                       \
        if (TRACE) {                                           \
            const long long _t0 = clock64();                   \
            stmt;                                              \
            tw[slot] += clock64() - _t0;                       \
        } else {                                               \
            stmt;                                              \
        }                                                      \
    } while (0)

template <bool TRACE>
__global__ void __launch_bounds__(192, 1) score_grad_duo_kernel(const __grid_constant__ CUtensorMap tmU,
                                                                const __grid_constant__ CUtensorMap tmE, GradParams P,
                                                                long long* __restrict__ trace) {
    long long tw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin = TRACE ? clock64() : 0;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ubuf = smem;
    uint8_t* ring = smem + D_UBYTES;
    uint8_t* pbuf = ring + D_STAGES * D_STAGE;
    DuoBarriers* bars = reinterpret_cast<DuoBarriers*>(pbuf + D_PBYTES);

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;     // (provably warp-uniform)
    const int m_tile = blockIdx.y, split = blockIdx.z;
    const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;       // cluster = (2, 1, 1): a CTA pair lies along x, rank == blockIdx.x
    const bool leader = rank == 0;
    const int n_tiles_total = (P.N + G_BN - 1) / G_BN;
    const int t0 = split * P.tiles_per_split, t1 = min(n_tiles_total, t0 + P.tiles_per_split);
    const int T = t1 - t0;
    const int KB = (P.d + BK - 1) / BK;        // 64-column k-blocks that hold data
    const int KS = (KB + 1) / 2;               // G1 ring stages per tile
    const int dpad = KS * 128;                 // O columns, a multiple of 128 (each CTA supplies half of every column group)
    const int NG = (dpad + 255) / 256;         // G2 column groups of 256 (the last one may be 128)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmU);
        tma_prefetch_desc(&tmE);
        for (int s = 0; s < D_STAGES; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->u_full, 1);
        mbar_init(&bars->o_full, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->s_full[a], 1); mbar_init(&bars->s_free[a], 8);      // 4 softmax warps x 2 CTAs
            mbar_init(&bars->p_full[a], 8); mbar_init(&bars->p_free[a], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair<512>(&bars->tmem_base);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // the peer's barriers are initialised before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_launch_dependents();       // only now: a dependent tensor-core kernel must not grab this SM's TMEM before we hold ours
    pdl_wait();
    const uint32_t tmem_O = tmem_base, tmem_S = tmem_base + D_TMEM_S;
    const bool onepass = P.c_part != nullptr;

    // softmax / epilogue threads: TMEM lane <-> (user, item half)
    const int q = warp & 3;
    const int tl = q * 32 + lane;              // TMEM lane of this thread
    const int u_local = tl & 63, half = tl >> 6;
    const int row = m_tile * BM + (int)rank * 64 + u_local;
    const bool row_ok = warp >= 2 && row < P.B;
    const uint32_t lane_base = uint32_t(q * 32) << 16;
    float lse_row = 0.f;           // two-pass: the row's log-sum-exp; one-pass: the shift
    if (row_ok && !onepass) {
        if (P.lse) lse_row = P.lse[row];
        else {
            lse_row = merged_lse(P.part_m, P.part_l, P.lse_splits, row);
            if (half == 0 && split == 0) P.lse_out[row] = lse_row;
        }
    }
    float lsum = 0.f;
    int rs = 0;                    // ring stage / phase of the producer or the MMA issuer; run on across sweeps
    uint32_t rph = 0;

    for (int sweep = 0;; ++sweep) {
        if (warp == 0) {
            // TMA producer.  The whole warp walks the loop and one elected lane issues (see the MMA warp).  Every load of either
            // CTA is credited to CTA 0's barrier; CTA 0 announces the bytes of both.  A ring stage is two [64 x 64] boxes:
            // with d % 64 == 0 ONE instruction fetches both through a 3-D view of the matrix (64 columns, rows, 64-column
            // blocks) -- at one 8 KB box per instruction the kernel was bound by TMA issue (~240 cycles per box on this thread).
            const uint32_t fbar0 = map_to_cta(smem_u32(&bars->full[0]), 0);
            auto load_stage = [&](const CUtensorMap* tm, int kblock0, int nb, int row0) {
                const int s = rs;
                DUO_TIMED(0, mbar_wait(&bars->empty[s], rph ^ 1));
                if (++rs == D_STAGES) { rs = 0; rph ^= 1; }
                if (elect_one()) {
                    uint8_t* dst = ring + s * D_STAGE;
                    if (P.box3d) {
                        if (leader) mbar_expect_tx(&bars->full[s], 2 * D_STAGE);
                        tma_load_3d_pair(dst, tm, 0, row0, kblock0, fbar0 + s * 8);
                    } else {
                        if (leader) mbar_expect_tx(&bars->full[s], 2 * nb * D_BOX);
                        for (int b = 0; b < nb; ++b) tma_load_2d_pair(dst + b * D_BOX, tm, (kblock0 + b) * BK, row0, fbar0 + s * 8);
                    }
                }
                __syncwarp();
            };
            if (sweep == 0) {      // my 64 user rows, resident for the whole kernel
                if (elect_one()) {
                    const uint32_t ubar = map_to_cta(smem_u32(&bars->u_full), 0);
                    const int row0 = m_tile * BM + (int)rank * 64;
                    if (P.box3d) {
                        if (leader) mbar_expect_tx(&bars->u_full, 2 * KS * D_STAGE);
                        for (int j = 0; j < KS; ++j) tma_load_3d_pair(ubuf + j * D_STAGE, &tmU, 0, row0, 2 * j, ubar);
                    } else {
                        if (leader) mbar_expect_tx(&bars->u_full, 2 * KB * D_BOX);
                        for (int kb = 0; kb < KB; ++kb) tma_load_2d_pair(ubuf + kb * D_BOX, &tmU, kb * BK, row0, ubar);
                    }
                }
                __syncwarp();
            }
            for (int ti = 0; ti <= T; ++ti) {
                if (ti < T) {          // G1 operands of tile ti: my 64 items, two k-blocks per stage
                    const int item0 = (t0 + ti) * G_BN + (int)rank * 64;
                    for (int j = 0; j < KS; ++j) load_stage(&tmE, 2 * j, min(2, KB - 2 * j), item0);
                }
                if (ti >= 1) {         // G2 operands of tile ti-1: both 64-item halves, my half of every column group
                    const int item0 = (t0 + ti - 1) * G_BN;
                    for (int ih = 0; ih < 2; ++ih)
                        for (int g = 0; g < NG; ++g) {
                            const int ncols = min(256, dpad - g * 256);
                            load_stage(&tmE, (g * 256 + (int)rank * (ncols / 2)) / BK, ncols / 128, item0 + ih * 64);
                        }
                }
            }
        } else if (warp == 1) {
            // The whole warp walks the loop (converged, warp-uniform values: the descriptor arithmetic stays in uniform
            // registers) and one elected lane issues -- an `if (lane == 0)` around it makes the compiler move every descriptor
            // into uniform registers through a per-MMA elect/broadcast loop, ~20 dependent instructions per MMA: too slow for
            // the 32-cycle 128x128x16 MMAs of G1.
            if (leader) {
                constexpr uint32_t idesc1 = make_idesc(BM, G_BN, 0, 0);
                const uint64_t dU = make_desc(smem_u32(ubuf), 16, 1024);               // + (byte offset >> 4)
                const uint64_t dRingK = make_desc(smem_u32(ring), 16, 1024);           // ring box read K-major (G1)
                const uint64_t dRingMN = make_desc(smem_u32(ring), D_BOX, 1024);       // ring boxes read MN-major (G2)
                const uint64_t dP = make_desc(smem_u32(pbuf), 16, 1024);
                if (sweep == 0) {
                    mbar_wait(&bars->u_full, 0);
                    tc_fence_after();
                }
                for (int ti = 0; ti <= T; ++ti) {
                    if (ti < T) {          // G1(ti): S = U E_t^T into S buffer b
                        const int g = sweep * T + ti, b = g & 1, k = g >> 1;
                        DUO_TIMED(0, mbar_wait(&bars->s_free[b], (k & 1) ^ 1));
                        tc_fence_after();
                        for (int j = 0; j < KS; ++j) {
                            const int s = rs;
                            DUO_TIMED(1, mbar_wait(&bars->full[s], rph));
                            if (++rs == D_STAGES) { rs = 0; rph ^= 1; }
                            tc_fence_after();
                            const int nb = min(2, KB - 2 * j);
                            if (elect_one()) {
                                const uint64_t da = dU + (uint32_t)((2 * j * D_BOX) >> 4), db = dRingK + (uint32_t)((s * D_STAGE) >> 4);
                                const uint32_t dst = tmem_S + b * 64;
#pragma unroll
                                for (int kk = 0; kk < BK / 16; ++kk) umma_ss_pair(dst, da + 2 * kk, db + 2 * kk, idesc1, (j | kk) != 0);
                                if (nb == 2) {
#pragma unroll
                                    for (int kk = 0; kk < BK / 16; ++kk)
                                        umma_ss_pair(dst, da + (D_BOX >> 4) + 2 * kk, db + (D_BOX >> 4) + 2 * kk, idesc1, 1);
                                }
                                umma_commit_pair(&bars->empty[s], 3);
                                if (j == KS - 1) umma_commit_pair(&bars->s_full[b], 3);
                            }
                            __syncwarp();
                        }
                    }
                    if (ti >= 1) {         // G2(ti-1): O += P E_t
                        const int tj = ti - 1, g = sweep * T + tj;
                        DUO_TIMED(2, mbar_wait(&bars->p_full[0], g & 1));      // both CTAs' softmax warps have written their P rows
                        tc_fence_after();
                        for (int ih = 0; ih < 2; ++ih)
                            for (int gq = 0; gq < NG; ++gq) {
                                const int ncols = min(256, dpad - gq * 256);
                                const uint32_t idesc2 = make_idesc(BM, ncols, 0, 1);
                                const int s = rs;
                                DUO_TIMED(3, mbar_wait(&bars->full[s], rph));
                                if (++rs == D_STAGES) { rs = 0; rph ^= 1; }
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint64_t da = dP + (uint32_t)((ih * D_BOX) >> 4);
                                    const uint64_t db = dRingMN + (uint32_t)((s * D_STAGE) >> 4);
                                    const uint32_t dst = tmem_O + gq * 128;
#pragma unroll
                                    for (int kk = 0; kk < 4; ++kk)      // K = 16 items per MMA: 32 bytes of a P row, 16 rows of the E box
                                        umma_ss_pair(dst, da + 2 * kk, db + (2048 >> 4) * kk, idesc2, (tj | ih | kk) != 0);
                                    umma_commit_pair(&bars->empty[s], 3);
                                    if (ih == 1 && gq == NG - 1) umma_commit_pair(&bars->p_free[0], 3);
                                }
                                __syncwarp();
                            }
                    }
                }
                if (elect_one()) umma_commit_pair(&bars->o_full, 3);
                __syncwarp();
            }
        } else {
            const float shift2 = lse_row * kLog2e;
            const uint32_t sfree0 = map_to_cta(smem_u32(&bars->s_free[0]), 0), pfull0 = map_to_cta(smem_u32(&bars->p_full[0]), 0);
            lsum = 0.f;
            for (int ti = 0; ti < T; ++ti) {
                const int g = sweep * T + ti, b = g & 1, k = g >> 1;
                DUO_TIMED(0, mbar_wait(&bars->s_full[b], k & 1));
                tc_fence_after();
                float v[2][32];
                tmem_ld32(tmem_S + b * 64 + lane_base, v[0]);
                tmem_ld32(tmem_S + b * 64 + lane_base + 32, v[1]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(sfree0 + b * 8);
                // ONE P buffer: G2(t-1), which reads it, is issued before G1(t+1), and this store happens while G1(t+1) runs
                DUO_TIMED(1, mbar_wait(&bars->p_free[0], (g & 1) ^ 1));
                // my 64 items = one 128-byte row of swizzle atom `half` of the P tile
                uint8_t* prow = pbuf + half * D_BOX + u_local * 128;
                const int n_valid = P.N - (t0 + ti) * G_BN - half * 64;      // of my 64 items (TMA zero-fills the rows beyond N)
                if (n_valid >= 64) lsum += duo_softmax_row<false>(v, shift2, 64, prow, u_local);
                else lsum += duo_softmax_row<true>(v, shift2, n_valid, prow, u_local);
                fence_proxy_async();           // my P row (own shared memory) is visible to the MMA's async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(pfull0);
            }
            DUO_TIMED(2, mbar_wait(&bars->o_full, sweep & 1));      // every MMA of the sweep has completed
            tc_fence_after();
        }
        __syncwarp();
        if (TRACE && lane == 0 && warp <= 2) {
            tw[7] = clock64() - t_begin;
            const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * 2 + rank;
            for (int i = 0; i < 8; ++i) trace[(cta * 3 + warp) * 8 + i] = tw[i];
        }
        if (!onepass) break;
        // the two item halves of a row live in two threads; the pair repeats a sweep together (its MMAs span both CTAs)
        if (warp >= 2) bars->rsum[half][u_local] = lsum;
        __syncthreads();
        const float both = warp >= 2 ? bars->rsum[0][u_local] + bars->rsum[1][u_local] : 1.0f;
        const bool over = both > kOnepassOver, under = both < kOnepassUnder;     // inf counts as over; NaN propagates
        const int mine = __syncthreads_or(over || under);
        if (threadIdx.x == 0) {
            bars->vote[rank] = (uint32_t)mine;
            st_cluster_u32(map_to_cta(smem_u32(&bars->vote[rank]), peer), (uint32_t)mine);
        }
        cluster_sync_all();
        const bool again = (bars->vote[0] | bars->vote[1]) != 0;
        if (!again || sweep + 1 >= kOnepassMaxSweeps) break;
        lse_row += over ? kOnepassRetry : under ? -kOnepassRetry : 0.f;
        cluster_sync_all();            // both CTAs have read this sweep's votes before the next sweep's overwrite them
    }
    if (warp >= 2) {
        // ---- O (TMEM) -> global partial: lanes 0..63 hold the first half of a column group, lanes 64..127 the second -------
        if (onepass && row_ok) {
            if (half == 0) P.c_part[(size_t)split * P.B + row] = lse_row;
            P.l_part[((size_t)split * 2 + half) * P.B + row] = lsum;
        }
        float* orow = P.Opart + ((size_t)split * P.B + row) * P.ldo;
        for (int g = 0; g < NG; ++g) {
            const int hc = min(256, dpad - g * 256) / 2;
            const int col0 = g * 256 + half * hc;
            for (int c = 0; c < hc / 32; ++c) {
                float v[32];
                tmem_ld32(tmem_O + lane_base + g * 128 + c * 32, v);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        if (col0 + c * 32 + j < P.ldo)
                            *reinterpret_cast<float4*>(orow + col0 + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // nobody leaves (or frees TMEM) while the pair's MMAs / remote arrives may still touch this CTA
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}
