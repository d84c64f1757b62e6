// Minimal CUDA-on-CPU shim: lets tests/emu/ compile convexadam_b200/csrc/*.cu with g++ (-DCA_EMU) and run each
// thread block as a set of fibers (ucontext) on one OS thread, blocks spread over OpenMP workers.
// TEST INFRASTRUCTURE ONLY -- used to debug kernel indexing/ordering in a container without a GPU.  Supports
// what the kernels use: blockIdx/threadIdx, static+dynamic __shared__, __syncthreads, warp shuffles, atomics,
// _rn intrinsics.  Floating point: build with -ffp-contract=off so a*b+c never fuses (matches explicit _rn calls).
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __shared__ static thread_local
#define __constant__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

namespace emu {
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void* dyn_smem();
void sync_block();
void sync_warp();
void yield();           // give the other fibers of the block a turn (spin-wait on shared state)
void* warp_scratch();   // 32 x 8 bytes for the calling thread's warp
int lane_id();
int warp_alive_mask_has(int lane);
}  // namespace emu

void ca_note_launch();
#define CA_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (ca_note_launch(), emu::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); }))
#define CA_LAUNCH_PDL CA_LAUNCH
static inline void ca_grid_dep_wait() {}
static inline void ca_grid_dep_launch() {}
#define CA_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::dyn_smem())
#define CA_DYN_SMEM_ALIGNED(T, name, al) T* name = reinterpret_cast<T*>(emu::dyn_smem())

static inline void __syncthreads() { emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }

// ---- arithmetic intrinsics (round-to-nearest; contraction disabled by the build flags) ----
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __int2float_rn(int a) { return (float)a; }

// IEEE half round trip (round-to-nearest-even), enough for ca_pack_field
static inline float emu_round_half(float f) {
    _Float16 h = (_Float16)f;
    return (float)h;
}

// ---- atomics ----
static inline unsigned long long atomicMin(unsigned long long* a, unsigned long long v) {
    unsigned long long old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(a, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
static inline double atomicAdd(double* a, double v) {
    unsigned long long* p = (unsigned long long*)a;
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED), nw;
    double o;
    do { memcpy(&o, &old, 8); double s = o + v; memcpy(&nw, &s, 8); } while (!__atomic_compare_exchange_n(p, &old, nw, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return o;
}
static inline float atomicAdd(float* a, float v) {
    unsigned* p = (unsigned*)a;
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED), nw;
    float o;
    do { memcpy(&o, &old, 4); float s = o + v; memcpy(&nw, &s, 4); } while (!__atomic_compare_exchange_n(p, &old, nw, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return o;
}

// ---- warp shuffles ----
template <class T> static inline T emu_shfl(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    char* s = (char*)emu::warp_scratch();
    memcpy(s + 8 * emu::lane_id(), &v, sizeof(T));
    emu::sync_warp();
    T r = v;
    if (src_lane >= 0 && src_lane < 32 && emu::warp_alive_mask_has(src_lane)) memcpy(&r, s + 8 * src_lane, sizeof(T));
    emu::sync_warp();
    return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, emu::lane_id() ^ m); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) { return emu_shfl(v, emu::lane_id() + d); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) { return emu_shfl(v, emu::lane_id() - d); }
template <class T> static inline T __shfl_sync(unsigned, T v, int l) { return emu_shfl(v, l); }
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) {
        const int p = emu_shfl(pred, l);
        if (emu::warp_alive_mask_has(l) && p) r |= 1u << l;
    }
    return r;
}
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
// product rounded towards +inf (the double product of two floats is exact)
static inline float __fmul_ru(float a, float b) {
    const double p = (double)a * (double)b;
    float r = (float)p;
    if ((double)r < p) r = nextafterf(r, INFINITY);
    return r;
}
