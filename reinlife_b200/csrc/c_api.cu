// c_api.cu -- error plumbing + version of libreinlife_b200.so (entry points live next to their kernels).
#include <stdarg.h>
#include "rl_common.cuh"

thread_local char g_rl_err[512] = "";

int rl_set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_rl_err, sizeof(g_rl_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" {
const char* rl_last_error(void) { return g_rl_err; }
int rl_version(void) { return 100; }
}
