// Batch transposition for the layer-1 weight gradient: the batch's (user, item, value) entries are sorted
// stably by item so that every touched row of W1^T is summed by one CTA in a fixed order (encoder.cu:
// w1_grad_kernel).  The sort / scan are CUB device primitives (plumbing, not a hot op: ~40k keys per step).
// This replaces the dense dW1 = dH^T . X GEMM of the reference's autograd.
#include <cub/cub.cuh>

#include "common.cuh"

namespace hvae {

__global__ void __launch_bounds__(1024) batch_offsets_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ rows, int B,
                                                              int32_t* __restrict__ boff) {
    pdl_prologue();
    typedef cub::BlockScan<int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ int carry;
    if (threadIdx.x == 0) { carry = 0; boff[0] = 0; }
    __syncthreads();
    for (int base = 0; base < B; base += 1024) {
        const int b = base + threadIdx.x;
        int len = 0;
        if (b < B) { const int u = rows ? rows[b] : b; len = u < 0 ? 0 : (int)(indptr[u + 1] - indptr[u]); }  // u < 0: padding slot
        int incl, total;
        Scan(tmp).InclusiveSum(len, incl, total);
        if (b < B) boff[b + 1] = carry + incl;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
}

__global__ void fill_keys_kernel(int32_t* __restrict__ keys, int32_t* __restrict__ eid, int cap, int sentinel) {
    pdl_prologue();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < cap) { keys[e] = sentinel; eid[e] = e; }
}

__global__ void expand_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
                              const int32_t* __restrict__ rows, int B, const int32_t* __restrict__ boff, int cap,
                              int32_t* __restrict__ keys, int32_t* __restrict__ ent_user, float* __restrict__ ent_val,
                              int32_t* __restrict__ overflow) {
    pdl_prologue();
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= B) return;
    const int u = rows ? rows[b] : b;
    if (u < 0) return;  // padding slot of a data-parallel global batch
    const int64_t s = indptr[u];
    const int len = (int)(indptr[u + 1] - s), o = boff[b];
    if (o + len > cap) { if (lane == 0) atomicAdd(overflow, 1); return; }   // sticky: hvae_batch_release poisons the step's loss
    for (int j = lane; j < len; j += 32) {
        keys[o + j] = indices[s + j];
        ent_user[o + j] = b;
        ent_val[o + j] = values ? values[s + j] : 1.0f;
    }
}

__global__ void seg_heads_kernel(const int32_t* __restrict__ keys, int cap, int sentinel, int32_t* __restrict__ head) {
    pdl_prologue();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cap) return;
    const int k = keys[e];
    head[e] = (k != sentinel && (e == 0 || k != keys[e - 1])) ? 1 : 0;
}

__global__ void seg_write_kernel(const int32_t* __restrict__ keys, const int32_t* __restrict__ head, const int32_t* __restrict__ slot,
                                 int cap, int sentinel, int32_t* __restrict__ seg_start, int32_t* __restrict__ uniq_item,
                                 int32_t* __restrict__ slot_of_item, int32_t* __restrict__ n_unique) {
    pdl_prologue();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cap) return;
    const int k = keys[e];
    if (head[e]) {
        const int s = slot[e];
        seg_start[s] = e;
        uniq_item[s] = k;
        slot_of_item[k] = s;
    }
    if (k == sentinel && (e == 0 || keys[e - 1] != sentinel)) { *n_unique = slot[e]; seg_start[slot[e]] = e; }
    if (e == cap - 1 && k != sentinel) { *n_unique = slot[e] + head[e]; seg_start[slot[e] + head[e]] = cap; }
}

// Also the step's overflow check: if hvae_batch_transpose had to drop rows (the caller's nnz bound was too small, so the
// layer-1 weight gradient of this step is incomplete) the step's loss scalars and the epoch accumulators become NaN -- the
// host raises on the next loss read (hvae_b200/train.py); the counter stays set until the host clears it.
__global__ void batch_release_kernel(const int32_t* __restrict__ uniq_item, const int32_t* __restrict__ n_unique, int cap,
                                     int32_t* __restrict__ slot_of_item, const int32_t* __restrict__ overflow,
                                     float* __restrict__ loss_out, float* __restrict__ acc) {
    pdl_prologue();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < cap && s < *n_unique) slot_of_item[uniq_item[s]] = -1;
    if (s == 0 && overflow && *overflow != 0) {
        const float qnan = __int_as_float(0x7fc00000);
        if (loss_out) { loss_out[0] = qnan; loss_out[1] = qnan; loss_out[2] = qnan; }
        if (acc) { acc[0] = qnan; acc[1] = qnan; acc[2] = qnan; }
    }
}

static int key_bits(int n_items) {
    int bits = 1;
    while ((1ll << bits) <= (long long)n_items) ++bits;
    return bits;
}

}  // namespace hvae

using namespace hvae;

extern "C" {

size_t hvae_batch_temp_bytes(int cap, int n_items) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const int32_t*)nullptr, (int32_t*)nullptr, (const int32_t*)nullptr, (int32_t*)nullptr,
                                    cap, 0, key_bits(n_items));
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, cap);
    return (a > b ? a : b) + 256;
}

int hvae_batch_offsets(const int64_t* indptr, const int32_t* rows, int B, int32_t* boff, void* stream) {
    launch_pdl(batch_offsets_kernel, 1, 1024, 0, (cudaStream_t)stream, indptr, rows, B, boff);
    HVAE_LAUNCH_CHECK("batch_offsets");
    return 0;
}

// int32 work arrays, each of `cap` entries unless noted: keys, keys_sorted, eid, eid_sorted, head, slot, ent_user, uniq_item;
// ent_val (float, cap); seg_start (cap+1); slot_of_item (n_items, must be all -1 on entry, restored by hvae_batch_release);
// n_unique, overflow: single int32 each.
int hvae_batch_transpose(const int64_t* indptr, const int32_t* indices, const float* values, const int32_t* rows, int B, int n_items,
                         int cap, const int32_t* boff, int32_t* keys, int32_t* keys_sorted, int32_t* eid, int32_t* eid_sorted,
                         int32_t* head, int32_t* slot, int32_t* ent_user, float* ent_val, int32_t* seg_start, int32_t* uniq_item,
                         int32_t* slot_of_item, int32_t* n_unique, int32_t* overflow, void* temp, size_t temp_bytes, void* stream) {
    HVAE_REQUIRE(cap >= 1, "batch_transpose: cap must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    const int tb = 256, gb = ceil_div(cap, tb);
    launch_pdl(fill_keys_kernel, gb, tb, 0, st, keys, eid, cap, n_items);
    if (B > 0)
        launch_pdl(expand_kernel, ceil_div(B, 8), 256, 0, st, indptr, indices, values, rows, B, boff, cap, keys, ent_user, ent_val, overflow);
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, keys, keys_sorted, eid, eid_sorted, cap, 0, key_bits(n_items), st);
    HVAE_REQUIRE(need <= temp_bytes, "batch_transpose: temp storage %zu < %zu", temp_bytes, need);
    HVAE_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_sorted, eid, eid_sorted, cap, 0, key_bits(n_items), st));
    launch_pdl(seg_heads_kernel, gb, tb, 0, st, keys_sorted, cap, n_items, head);
    HVAE_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, head, slot, cap, st));
    launch_pdl(seg_write_kernel, gb, tb, 0, st, keys_sorted, head, slot, cap, n_items, seg_start, uniq_item, slot_of_item, n_unique);
    HVAE_LAUNCH_CHECK("batch_transpose");
    return 0;
}

int hvae_batch_release(const int32_t* uniq_item, const int32_t* n_unique, int cap, int32_t* slot_of_item, const int32_t* overflow,
                       float* loss_out, float* acc, void* stream) {
    launch_pdl(batch_release_kernel, ceil_div(cap, 256), 256, 0, (cudaStream_t)stream, uniq_item, n_unique, cap, slot_of_item, overflow,
               loss_out, acc);
    HVAE_LAUNCH_CHECK("batch_release");
    return 0;
}

}  // extern "C"
