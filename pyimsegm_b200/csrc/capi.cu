// capi.cu -- error plumbing of the C-ABI (include/imsegm_b200.h)
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";
long long g_isb_launches = 0;

void isb_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* isb_last_error(void) { return g_err; }
extern "C" int isb_abi_version(void) { return 1; }
extern "C" long long isb_launch_count(void) { return g_isb_launches; }
