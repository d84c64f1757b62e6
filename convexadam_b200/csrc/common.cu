// Error plumbing + library identity of libconvexadam_b200.
#include "ca_platform.h"

#include <cstdarg>
#include <mutex>

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

// ---- optional per-kernel timing (bench.py's roofline line) ------------------------------------------------------------
// When enabled, the launches of a few named kernels are bracketed by CUDA events recorded on the launch stream itself
// (no synchronisation; the events are read back by ca_kernel_times after the caller has synchronised).
#ifndef CA_EMU
namespace {
struct TimedLaunch { cudaEvent_t e0, e1; int slot; };
constexpr int kMaxTimed = 4096;
TimedLaunch g_timed[kMaxTimed];
int g_ntimed = 0, g_ncreated = 0;
unsigned g_timing = 0;      // bit mask of the slots to time
std::mutex g_timing_mu;     // the slot table is shared by every host thread that launches through the library
}  // namespace
void* ca_timing_begin(int slot, cudaStream_t st) {
    if (!((__atomic_load_n(&g_timing, __ATOMIC_RELAXED) >> slot) & 1u)) return nullptr;     // the common case: timing off
    std::lock_guard<std::mutex> lock(g_timing_mu);
    if (g_ntimed >= kMaxTimed) return nullptr;
    TimedLaunch& t = g_timed[g_ntimed];
    if (g_ntimed >= g_ncreated) {
        if (cudaEventCreate(&t.e0) != cudaSuccess || cudaEventCreate(&t.e1) != cudaSuccess) return nullptr;
        ++g_ncreated;
    }
    t.slot = slot;
    cudaEventRecord(t.e0, st);
    ++g_ntimed;
    return &t;
}
void ca_timing_end(void* h, cudaStream_t st) {
    if (h) cudaEventRecord(((TimedLaunch*)h)->e1, st);
}
extern "C" void ca_kernel_timing(int slot_mask) {
    std::lock_guard<std::mutex> lock(g_timing_mu);
    __atomic_store_n(&g_timing, (unsigned)slot_mask, __ATOMIC_RELAXED);
    g_ntimed = 0;
}
// sums the recorded launches per slot (ms[slot], count[slot], slot < nslots); the caller synchronises first
extern "C" int ca_kernel_times(double* ms, int* count, int nslots) {
    std::lock_guard<std::mutex> lock(g_timing_mu);
    for (int i = 0; i < nslots; ++i) { ms[i] = 0.0; count[i] = 0; }
    for (int i = 0; i < g_ntimed; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, g_timed[i].e0, g_timed[i].e1) != cudaSuccess) return CA_ERR_LAUNCH;
        if (g_timed[i].slot < nslots) { ms[g_timed[i].slot] += t; count[g_timed[i].slot] += 1; }
    }
    return g_ntimed;
}
#else
void* ca_timing_begin(int, cudaStream_t) { return nullptr; }
void ca_timing_end(void*, cudaStream_t) {}
extern "C" void ca_kernel_timing(int) {}
extern "C" int ca_kernel_times(double*, int*, int) { return 0; }
#endif
