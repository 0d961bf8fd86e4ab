// ugs_capi.cu -- error plumbing and ABI version of libugs_b200.so (entry points live next to their kernels).
#include <stdarg.h>
#include <stdio.h>
#include "ugs_common.cuh"

static thread_local char g_err[512] = "";

int ugs_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* ugs_last_error(void) { return g_err; }
extern "C" int ugs_abi_version(void) { return UGS_B200_ABI_VERSION; }

unsigned long long g_ugs_launches = 0;
extern "C" uint64_t ugs_launch_count(void) { return g_ugs_launches; }
