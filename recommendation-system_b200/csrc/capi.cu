// Error channel and version of the C ABI (include/hvae_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "hvae_b200.h"

static thread_local char g_err[512] = "";

int hvae_fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

bool hvae::pdl_enabled() {
    static const bool on = getenv("HVAE_NO_PDL") == nullptr;
    return on;
}

extern "C" const char* hvae_last_error(void) { return g_err; }
extern "C" int hvae_abi_version(void) { return 2; }

// ---- host <-> device plumbing of the per-step API (VAETrainer.train_on_batch) --------------------------------------
// The batch's CSR slice from (pinned) host memory into the step's static device buffers as three async copies behind one
// call, and the step's loss scalars back (async copy + wait for the stream): per step this replaces ~10 framework calls.
extern "C" int hvae_h2d_csr_batch(const int64_t* h_crow, const int32_t* h_col, const float* h_val, int B, int64_t nnz,
                                  int64_t* d_crow, int32_t* d_col, float* d_val, void* stream) {
    HVAE_REQUIRE(B >= 0 && nnz >= 0 && h_crow && d_crow, "h2d_csr_batch: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    HVAE_CUDA(cudaMemcpyAsync(d_crow, h_crow, sizeof(int64_t) * (size_t)(B + 1), cudaMemcpyHostToDevice, st));
    if (nnz > 0) {
        HVAE_REQUIRE(h_col && d_col && h_val && d_val, "h2d_csr_batch: missing column / value arrays");
        HVAE_CUDA(cudaMemcpyAsync(d_col, h_col, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
        HVAE_CUDA(cudaMemcpyAsync(d_val, h_val, sizeof(float) * (size_t)nnz, cudaMemcpyHostToDevice, st));
    }
    return 0;
}

extern "C" int hvae_d2h_floats(const float* d_src, int n, float* h_dst, void* stream) {
    HVAE_REQUIRE(n >= 0 && (n == 0 || (d_src && h_dst)), "d2h_floats: bad arguments");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    HVAE_CUDA(cudaMemcpyAsync(h_dst, d_src, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, st));
    HVAE_CUDA(cudaStreamSynchronize(st));
    return 0;
}
