// DRAFT for the next round -- NOT part of libhvae_b200.so (build.py compiles csrc/*.cu only), never run on a GPU yet.
// Compile check:  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 --expt-relaxed-constexpr -I include
//                      -I recommendation-system_b200/csrc -c recommendation-system_b200/csrc/next/score_ksplit.cu -o /dev/null
//
// K-split variant of the one-pass scoring kernel for 384 < d <= 768.  The two CTAs of a cluster split the
// d axis: CTA r owns columns [384 r, 384 r + 384) of U (RESIDENT in shared memory), of every E tile and of O.  A 64-item
// tile of E is loaded ONCE per CTA (six [64 items x 64 cols] boxes) and serves both GEMMs:
//     G1  S_r = U_r E_{t,r}^T      (boxes read K-major,  N = 64, accumulators: 64 TMEM columns, double-buffered)
//     G2  O_r += P E_{t,r}         (the same boxes read MN-major, K = 64 items)
// S = S_0 + S_1 takes one DSMEM exchange per tile: thread <-> user row keeps the partial scores of "its" 32 items, sends the
// other 32 (fp32) into the peer's exchange buffer, adds what it receives, takes exp2 against the one-pass shift and writes its
// 32 bf16 numerators into BOTH CTAs' P tile.  L2->SM traffic per CTA and 128 items: 96 KiB (pair kernel: 288 KiB).
//
// Shared memory: U 96 KiB | E tiles 2 x 48 KiB | P 16 KiB | X 16 KiB | barriers.   TMEM: O 384 + S 2 x 64 columns.
#include <cstdlib>

#include "common.cuh"
#include "hvae_b200.h"
#include "tc_common.cuh"

namespace hvae {
namespace tc {
namespace next {

constexpr int BM = 128, BK = 64;
constexpr int KS_BN = 64;                         // items per tile
constexpr int KS_HALF = 384, KS_KB = KS_HALF / BK;   // columns / k-blocks per CTA
constexpr int KS_UBLK = BM * BK * 2;              // 16 KiB: one resident U k-block [128 users x 64 cols]
constexpr int KS_BOX = KS_BN * BK * 2;            // 8 KiB: [64 items x 64 cols]
constexpr int KS_TILE = KS_KB * KS_BOX;           // 48 KiB: one E tile of this CTA's columns
constexpr int KS_PBYTES = BM * KS_BN * 2;         // 16 KiB: numerators [128 users x 64 items] bf16, 128B-swizzled rows
constexpr int KS_XBYTES = BM * 32 * 4;            // 16 KiB: the peer's partial scores of my 32 items, [row][32] fp32
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kRetry = 60.0f, kOver = 1.2676506e30f /* 2^100 */, kUnder = 8.8817842e-16f /* 2^-50 */;
constexpr int kMaxSweeps = 64;

struct KsParams {
    int B, N, d;
    int tiles_per_split, n_splits;     // 64-item tiles
    const float* lse;                  // two-pass mode: the rows' log-sum-exp (one-pass: null)
    float* Opart;                      // [n_splits][B][ldo]
    int ldo;
    float* c_part;                     // one-pass mode: [n_splits][B]
    float* l_part;                     //                [n_splits][2][B]
};

struct __align__(8) KsBarriers {
    uint64_t u_full;
    uint64_t full[2][KS_KB];           // box j of tile slot s has landed
    uint64_t empty[2][2];              // column group g (boxes 0-3 / 4-5) of tile slot s has been consumed by G2
    uint64_t s_full[2], s_free[2];     // S accumulator buffers
    uint64_t x_full;                   // the peer's partial scores are in my X buffer      (4 remote warp arrivals)
    uint64_t x_free;                   // the peer has consumed what I wrote into ITS X     (4 remote warp arrivals)
    uint64_t p_full;                   // both halves of P are in my P buffer              (4 local + 4 remote warp arrivals)
    uint64_t p_free_local;             // my G2 has consumed my P buffer                    (tcgen05.commit, local)
    uint64_t p_free_peer;              // the peer's G2 has consumed ITS P buffer           (tcgen05.commit from the peer)
    uint64_t o_full;
    uint32_t tmem_base;
    float rsum[2][BM];                 // one-pass: the row sums of both CTAs' sweeps
};

constexpr size_t kKsSmem = KS_KB * KS_UBLK + 2 * KS_TILE + KS_PBYTES + KS_XBYTES + 2048 + 1024;
static_assert(kKsSmem <= 232448, "shared memory budget");
static_assert(sizeof(KsBarriers) <= 2048, "barrier block");

__global__ void __launch_bounds__(192, 1) score_grad_ksplit_kernel(const __grid_constant__ CUtensorMap tmU,
                                                                   const __grid_constant__ CUtensorMap tmE, KsParams P) {
    pdl_launch_dependents();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ures = smem;                                   // 6 x 16 KiB
    uint8_t* etile = ures + KS_KB * KS_UBLK;                // 2 x 48 KiB
    uint8_t* pbuf = etile + 2 * KS_TILE;                    // 16 KiB
    float* xbuf = reinterpret_cast<float*>(pbuf + KS_PBYTES);   // 16 KiB
    KsBarriers* bars = reinterpret_cast<KsBarriers*>(pbuf + KS_PBYTES + KS_XBYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, split = blockIdx.z;
    const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;       // cluster of 2 along y: rank == K half
    const int n_tiles_total = (P.N + KS_BN - 1) / KS_BN;
    const int t0 = split * P.tiles_per_split, t1 = min(n_tiles_total, t0 + P.tiles_per_split);
    const int T = t1 - t0;
    const int dpad = (P.d + BK - 1) / BK * BK;
    const int dc0 = (int)rank * KS_HALF;                              // first column of this CTA
    const int KBh = min(KS_KB, (dpad - dc0) / BK);                    // k-blocks (= 64-column boxes) of this CTA: 1..6
    const int DC = KBh * BK;
    const int NG = (DC + 255) / 256;                                  // G2 column groups (<= 256 columns per MMA)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmU);
        tma_prefetch_desc(&tmE);
        mbar_init(&bars->u_full, 1);
        for (int s = 0; s < 2; ++s) {
            for (int j = 0; j < KS_KB; ++j) mbar_init(&bars->full[s][j], 1);
            for (int g = 0; g < 2; ++g) mbar_init(&bars->empty[s][g], 1);
            mbar_init(&bars->s_full[s], 1);
            mbar_init(&bars->s_free[s], 4);
        }
        mbar_init(&bars->x_full, 4); mbar_init(&bars->x_free, 4);
        mbar_init(&bars->p_full, 8); mbar_init(&bars->p_free_local, 1); mbar_init(&bars->p_free_peer, 1);
        mbar_init(&bars->o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&bars->tmem_base);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // the peer's barriers exist before anybody arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();
    const uint32_t tmem_O = tmem_base, tmem_S = tmem_base + KS_HALF;   // S buffer b at tmem_S + 64 b
    const bool onepass = P.c_part != nullptr;

    const int q = warp & 3;
    const int r_local = q * 32 + lane;
    const int row = m_tile * BM + r_local;
    const bool row_ok = warp >= 2 && row < P.B;
    const uint32_t lane_base = uint32_t(q * 32) << 16;
    float shift = (row_ok && !onepass) ? P.lse[row] : 0.f;
    float lsum = 0.f;

    // resident half of U: once per CTA
    if (warp == 0 && lane == 0) {
        mbar_expect_tx(&bars->u_full, (uint32_t)(KBh * KS_UBLK));
        for (int j = 0; j < KBh; ++j) tma_load_2d(ures + j * KS_UBLK, &tmU, dc0 + j * BK, m_tile * BM, &bars->u_full);
    }

    for (int sweep = 0;; ++sweep) {
        if (warp == 0) {
            if (lane == 0) {
                for (int t = 0; t < T; ++t) {
                    const int g = sweep * T + t, s = g & 1, use = g >> 1;       // use-th occupancy of tile slot s
                    const int item0 = (t0 + t) * KS_BN;
                    for (int j = 0; j < KBh; ++j) {
                        if (j == 0 || j == 4) mbar_wait(&bars->empty[s][j >> 2], (use & 1) ^ 1);     // column group free again
                        mbar_expect_tx(&bars->full[s][j], KS_BOX);
                        tma_load_2d(etile + s * KS_TILE + j * KS_BOX, &tmE, dc0 + j * BK, item0, &bars->full[s][j]);
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                constexpr uint32_t idesc1 = make_idesc(BM, KS_BN, 0, 0);
                if (sweep == 0) { mbar_wait(&bars->u_full, 0); tc_fence_after(); }
                for (int ti = 0; ti <= T; ++ti) {
                    if (ti < T) {          // G1(ti): S_r = U_r E_{t,r}^T  -- the boxes stay for G2(ti)
                        const int g = sweep * T + ti, s = g & 1, use = g >> 1;
                        mbar_wait(&bars->s_free[s], (use & 1) ^ 1);
                        tc_fence_after();
                        for (int j = 0; j < KBh; ++j) {
                            mbar_wait(&bars->full[s][j], use & 1);
                            tc_fence_after();
                            const uint32_t a0 = smem_u32(ures + j * KS_UBLK), b0 = smem_u32(etile + s * KS_TILE + j * KS_BOX);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k)
                                umma_ss(tmem_S + s * KS_BN, make_desc(a0 + k * 32, 16, 1024), make_desc(b0 + k * 32, 16, 1024), idesc1, (j | k) != 0);
                        }
                        umma_commit(&bars->s_full[s]);
                    }
                    if (ti >= 1) {         // G2(ti-1): O_r += P E_{t,r}
                        const int tj = ti - 1, g = sweep * T + tj, s = g & 1;
                        mbar_wait_cluster(&bars->p_full, g & 1);          // half of P was written by the peer
                        fence_proxy_async_all();
                        tc_fence_after();
                        const uint32_t p0 = smem_u32(pbuf);
                        for (int gq = 0; gq < NG; ++gq) {
                            const int ncols = min(256, DC - gq * 256);
                            const uint32_t idesc2 = make_idesc(BM, ncols, 0, 1);
                            const uint32_t b0 = smem_u32(etile + s * KS_TILE + gq * 4 * KS_BOX);
#pragma unroll
                            for (int kk = 0; kk < KS_BN / 16; ++kk)       // K = 16 items per MMA
                                umma_ss(tmem_O + gq * 256, make_desc(p0 + kk * 32, 16, 1024), make_desc(b0 + kk * 2048, KS_BOX, 1024), idesc2,
                                        (tj | kk) != 0);
                            umma_commit(&bars->empty[s][gq]);
                        }
                        if (NG == 1) umma_commit(&bars->empty[s][1]);      // (group 1 does not exist: keep its phase in step)
                        umma_commit(&bars->p_free_local);
                        umma_commit_cluster(map_to_cta(smem_u32(&bars->p_free_peer), peer));
                    }
                }
                umma_commit(&bars->o_full);
            }
        } else {
            const float shift2 = shift * kLog2e;
            const int h = (int)rank;                                             // my half of a tile's items: [32 h, 32 h + 32)
            const uint32_t xrow_peer = map_to_cta(smem_u32(xbuf + r_local * 32), peer);
            const uint32_t prow_local = smem_u32(pbuf + r_local * 128);
            const uint32_t prow_peer = map_to_cta(prow_local, peer);
            const uint32_t xfull_peer = map_to_cta(smem_u32(&bars->x_full), peer);
            const uint32_t xfree_peer = map_to_cta(smem_u32(&bars->x_free), peer);
            const uint32_t pfull_peer = map_to_cta(smem_u32(&bars->p_full), peer);
            lsum = 0.f;
            for (int t = 0; t < T; ++t) {
                const int g = sweep * T + t, s = g & 1, use = g >> 1;
                mbar_wait(&bars->s_full[s], use & 1);
                tc_fence_after();
                float lo[32], hi[32], mine[32], theirs[32];
                tmem_ld32(tmem_S + lane_base + s * KS_BN, lo);
                tmem_ld32(tmem_S + lane_base + s * KS_BN + 32, hi);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) { mine[i] = h ? hi[i] : lo[i]; theirs[i] = h ? lo[i] : hi[i]; }   // (selects: no dynamic register indexing)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->s_free[s]);
                // ---- exchange: the peer's 32 items to the peer, its partials of my 32 items to me -------------------
                mbar_wait_cluster(&bars->x_free, (g & 1) ^ 1);               // the peer has read my previous send
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    st_cluster_v4(xrow_peer + c * 16, __float_as_uint(theirs[4 * c]), __float_as_uint(theirs[4 * c + 1]),
                                  __float_as_uint(theirs[4 * c + 2]), __float_as_uint(theirs[4 * c + 3]));
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(xfull_peer);              // release.cluster
                mbar_wait_cluster(&bars->x_full, g & 1);
                float sc[32];
                const float4* xr = reinterpret_cast<const float4*>(xbuf + r_local * 32);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 x = xr[c];
                    sc[4 * c] = mine[4 * c] + x.x; sc[4 * c + 1] = mine[4 * c + 1] + x.y;
                    sc[4 * c + 2] = mine[4 * c + 2] + x.z; sc[4 * c + 3] = mine[4 * c + 3] + x.w;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(xfree_peer);              // my X buffer may be overwritten
                // ---- numerators of my 32 items -> both CTAs' P tile --------------------------------------------------
                const int n_valid = P.N - (t0 + t) * KS_BN - 32 * h;         // items of my half inside the catalogue
                mbar_wait(&bars->p_free_local, (g & 1) ^ 1);
                mbar_wait_cluster(&bars->p_free_peer, (g & 1) ^ 1);
#pragma unroll
                for (int c = 0; c < 4; ++c) {                                // 8 items -> one 16-byte chunk
                    uint32_t w[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = 8 * c + 2 * e;
                        float p0 = exp2f(fmaf(sc[i], kLog2e, -shift2)), p1 = exp2f(fmaf(sc[i + 1], kLog2e, -shift2));
                        if (i >= n_valid) p0 = 0.f;
                        if (i + 1 >= n_valid) p1 = 0.f;
                        lsum += p0 + p1;
                        __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
                        w[e] = *reinterpret_cast<uint32_t*>(&hh);
                    }
                    const uint32_t off = (uint32_t)((((4 * h + c) ^ (r_local & 7)) << 4));      // chunk 4h+c of the row, 128B swizzle
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow_local + off), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                    st_cluster_v4(prow_peer + off, w[0], w[1], w[2], w[3]);
                }
                fence_proxy_async_all();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bars->p_full);
                    mbar_arrive_cluster(pfull_peer);
                }
            }
            mbar_wait(&bars->o_full, sweep & 1);
            tc_fence_after();
        }
        __syncwarp();
        if (!onepass) break;
        if (warp >= 2) {           // the sweep's row sums of both CTAs decide together (each summed its own 32 items per tile)
            bars->rsum[rank][r_local] = lsum;
            const uint32_t remote = map_to_cta(smem_u32(&bars->rsum[rank][r_local]), peer);
            asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(lsum) : "memory");
        }
        cluster_sync_all();
        const float both = warp >= 2 ? bars->rsum[0][r_local] + bars->rsum[1][r_local] : 1.0f;
        const bool over = both > kOver, under = both < kUnder;
        if (!__syncthreads_or(over || under) || sweep + 1 >= kMaxSweeps) break;
        shift += over ? kRetry : under ? -kRetry : 0.f;
        cluster_sync_all();
    }
    if (warp >= 2) {
        if (onepass && row_ok) {
            if (rank == 0) P.c_part[(size_t)split * P.B + row] = shift;
            P.l_part[((size_t)split * 2 + rank) * P.B + row] = lsum;
        }
        float* orow = P.Opart + ((size_t)split * P.B + row) * P.ldo + dc0;
        for (int c = 0; c < DC / 32; ++c) {
            float v[32];
            tmem_ld32(tmem_O + lane_base + c * 32, v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (dc0 + c * 32 + j < P.ldo) *reinterpret_cast<float4*>(orow + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// Launch (to be added to launch_grad in score_tc.cu once validated):  grid (ceil(B/128), 2, n_splits), cluster (1, 2, 1),
// 192 threads, kKsSmem dynamic shared memory;  tmU = make_tmap_bf16(U, B, d, ldu, 128), tmE = make_tmap_bf16(E, N, d, lde, 64);
// n_tiles = ceil(N / 64), n_splits = pick_wave_splits(2 * m_tiles, n_tiles);  outputs exactly as score_grad_pair_kernel
// (Opart, c_part [S][B], l_part [S][2][B]), so hvae_tc_onepass_combine / hvae_du_finalize apply unchanged.
// Validation plan: tests/test_gpu_tc.py::test_tc_onepass_* with HVAE_KSPLIT=1 (d = 448, 768; ragged B / N; sweep repeats).

}  // namespace next
}  // namespace tc
}  // namespace hvae
