// Error plumbing + library identity of libconvexadam_b200.
#include "ca_platform.h"

#include <cstdarg>

static thread_local char g_ca_error[512] = "";

void ca_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_ca_error, sizeof(g_ca_error), fmt, ap);
    va_end(ap);
}

extern "C" int ca_abi_version(void) { return CA_ABI_VERSION; }
extern "C" const char* ca_last_error(void) { return g_ca_error; }
extern "C" int ca_is_cuda_build(void) {
#ifdef CA_EMU
    return 0;
#else
    return 1;
#endif
}
