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

static unsigned long long g_ca_launches = 0;
void ca_note_launch() { __atomic_fetch_add(&g_ca_launches, 1ull, __ATOMIC_RELAXED); }
extern "C" unsigned long long ca_launch_count(void) { return __atomic_load_n(&g_ca_launches, __ATOMIC_RELAXED); }

extern "C" int ca_abi_version(void) { return CA_ABI_VERSION; }
extern "C" const char* ca_last_error(void) { return g_ca_error; }
extern "C" int ca_is_cuda_build(void) {
#ifdef CA_EMU
    return 0;
#else
    return 1;
#endif
}
