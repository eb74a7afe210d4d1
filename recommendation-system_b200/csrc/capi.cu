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
extern "C" int hvae_abi_version(void) { return 1; }
