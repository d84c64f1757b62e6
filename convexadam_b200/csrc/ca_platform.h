// Platform glue for the convexadam_b200 kernels.
//
// Product build: nvcc, sm_100a (see __graft_entry__.build()).  The CA_EMU branch exists only so that
// tests/emu/ can compile the *same kernel sources* with g++ and run them block-by-block on the CPU
// (fibers = CUDA threads) to debug indexing without a GPU.  The product library is never built with
// CA_EMU and the Python package refuses to load anything but the CUDA build.
#pragma once

#ifdef CA_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CA_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (ca_note_launch(), kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__))
// Programmatic dependent launch: the kernel is queued behind its predecessor on the stream but its CTAs may be scheduled
// while the predecessor drains; it must call ca_grid_dep_wait() before touching anything the predecessor wrote (or reads).
// Chains of short dependent kernels (the Adam loop: 5 launches of 8-50 us per iteration) lose their launch gaps this way.
template <class... KArgs, class... Args>
static inline cudaError_t ca_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define CA_LAUNCH_PDL(kernel, grid, block, smem, stream, ...) \
    (ca_note_launch(), (void)ca_launch_pdl(kernel, (grid), (block), (smem), (stream), __VA_ARGS__))
static __device__ __forceinline__ void ca_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
static __device__ __forceinline__ void ca_grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define CA_DYN_SMEM(T, name)                                        \
    extern __shared__ __align__(16) unsigned char ca_dyn_smem_raw[]; \
    T* name = reinterpret_cast<T*>(ca_dyn_smem_raw)
#define CA_DYN_SMEM_ALIGNED(T, name, al)                                  \
    extern __shared__ __align__(al) unsigned char ca_dyn_smem_raw_al[];   \
    T* name = reinterpret_cast<T*>(ca_dyn_smem_raw_al)
#endif

#include "../../include/convexadam_b200.h"

void ca_note_launch();   // counts kernel launches (ca_launch_count) -- the bench reports it
// optional per-kernel timing with CUDA events on the launch stream (common.cu); slots:
enum { CA_T_RAW = 0, CA_T_BOX = 1, CA_T_BLOCKMIN = 2, CA_T_SELECT = 3, CA_T_FIXUP = 4, CA_T_ADAM_GRAD = 5, CA_T_BOX3 = 6, CA_T_FUSED = 7, CA_T_PIPE = 8, CA_T_NSLOTS = 9 };
void* ca_timing_begin(int slot, cudaStream_t st);
void ca_timing_end(void* h, cudaStream_t st);

// ---- error plumbing (no exceptions across the C ABI) -------------------------------------------------
void ca_set_error(const char* fmt, ...);
#define CA_CHECK_ARG(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            ca_set_error(__VA_ARGS__);   \
            return CA_ERR_BAD_ARGUMENT;  \
        }                                \
    } while (0)
// Opt a kernel in to > 48 KB of dynamic shared memory.  The attribute belongs to the (function, device) pair, so it is
// set before every such launch on whichever device is current (a host-side table lookup in the runtime; CUTLASS does the
// same) -- a per-process "done" flag would leave every device but the first without it.
#ifdef CA_EMU
#define CA_SMEM_OPTIN(kernel, bytes) do { } while (0)
#else
#define CA_SMEM_OPTIN(kernel, bytes)                                                                              \
    do {                                                                                                          \
        cudaError_t e__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
        if (e__ != cudaSuccess) {                                                                                 \
            ca_set_error("%s: cannot opt in to %d bytes of shared memory: %s", #kernel, (int)(bytes), cudaGetErrorString(e__)); \
            return CA_ERR_LAUNCH;                                                                                 \
        }                                                                                                         \
    } while (0)
#endif
#define CA_CHECK_LAUNCH(what)                                                       \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            ca_set_error("%s: launch failed: %s", what, cudaGetErrorString(e__));   \
            return CA_ERR_LAUNCH;                                                   \
        }                                                                           \
    } while (0)

// ---- small device helpers ---------------------------------------------------------------------------
static __host__ __device__ __forceinline__ int ca_min(int a, int b) { return a < b ? a : b; }
static __host__ __device__ __forceinline__ int ca_max(int a, int b) { return a > b ? a : b; }
static __host__ __device__ __forceinline__ int ca_clamp(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static __host__ __device__ __forceinline__ int64_t ca_divup(int64_t a, int64_t b) { return (a + b - 1) / b; }
static __host__ __device__ __forceinline__ int ca_roundup4(int a) { return (a + 3) & ~3; }

// x / 27 etc. with IEEE round-to-nearest, never contracted with neighbours (torch: `sum / divide_factor`).
static __device__ __forceinline__ float ca_div(float a, float b) { return __fdiv_rn(a, b); }

// x / 27 for the 3x3x3 box filters: q0 = RN(x*c), r = fma(-27, q0, x) (exact), q = fma(r, c, q0), c = RN(1/27).
// Bit-identical to the IEEE division for EVERY finite fp32 x (exhaustively verified over all 2^31 non-negative
// floats incl. subnormals by tests/div27_exhaustive.c; the ops are sign-symmetric) at 3 instructions instead of the
// ~12 of __fdiv_rn -- the box filters are issue-bound, so this matters.  (-0 maps to +0: never compared.)
static __device__ __forceinline__ float ca_div27(float x) {
    const float c = 1.0f / 27.0f;
    const float q0 = __fmul_rn(x, c);
    const float r = __fmaf_rn(-27.0f, q0, x);
    return __fmaf_rn(r, c, q0);
}

// streaming (read-once) 128-bit load: do not pollute L1
static __device__ __forceinline__ float4 ca_ld_stream4(const float4* p) {
#if defined(__CUDA_ARCH__)
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
#else
    return *p;
#endif
}
static __device__ __forceinline__ float2 ca_ld_stream2(const float2* p) {
#if defined(__CUDA_ARCH__)
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
#else
    return *p;
#endif
}
static __device__ __forceinline__ float ca_ld_stream1(const float* p) {
#if defined(__CUDA_ARCH__)
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
#else
    return *p;
#endif
}
// 128-bit read-only loads with an L2 eviction priority (createpolicy + ld.global.nc.L2::cache_hint): `first` for data read
// once per pass (it must not push the re-used working set out of L2), `last` for the working set itself
#if defined(__CUDA_ARCH__)
static __device__ __forceinline__ unsigned long long ca_l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
static __device__ __forceinline__ unsigned long long ca_l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
static __device__ __forceinline__ float4 ca_ld_hint4(const float4* ptr, unsigned long long policy) {
    float4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(ptr), "l"(policy));
    return r;
}
#else
static inline unsigned long long ca_l2_policy_evict_first() { return 0; }
static inline unsigned long long ca_l2_policy_evict_last() { return 0; }
static inline float4 ca_ld_hint4(const float4* ptr, unsigned long long) { return *ptr; }
#endif

// streaming (write-once) 128-bit store: evict-first, the cost volume must not push the working set out of L2
static __device__ __forceinline__ void ca_st_stream4(float4* p, const float4& v) {
#if defined(__CUDA_ARCH__)
    __stcs(p, v);
#else
    *p = v;
#endif
}

// ---- bulk asynchronous copies global -> shared memory (TMA engine, no tensor map: contiguous runs) --------------------
// One mbarrier per CTA phase: thread 0 arms it with the total byte count, any thread issues copies that complete on it,
// every thread waits for the phase.  Under CA_EMU the copy is a memcpy and the wait is a block barrier.
#if defined(CA_EMU)
static inline void ca_mbar_init(unsigned long long* bar, int count) { (void)bar; (void)count; }
static inline void ca_mbar_expect_tx(unsigned long long* bar, unsigned bytes) { (void)bar; (void)bytes; }
static inline void ca_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) { (void)bar; memcpy(dst, src, bytes); }
static inline void ca_mbar_wait(unsigned long long* bar, unsigned parity) { (void)bar; (void)parity; emu::sync_block(); }
#else
static __device__ __forceinline__ unsigned ca_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void ca_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ca_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
static __device__ __forceinline__ void ca_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ca_smem_u32(bar)), "r"(bytes) : "memory");
}
// bytes: a multiple of 16; dst and src 16-byte aligned
static __device__ __forceinline__ void ca_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ca_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(ca_smem_u32(bar))
                 : "memory");
}
static __device__ __forceinline__ void ca_mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CA_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CA_MBAR_DONE;\n"
        "bra CA_MBAR_WAIT;\n"
        "CA_MBAR_DONE:\n"
        "}\n" ::"r"(ca_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#endif

// ---- mbarriers as producer/consumer hand-offs between warp groups (arrive / wait by different threads) ------------------
// ca_mbar_* above are only ever used as "everybody waits for a copy" (CA_EMU: a block barrier); these carry the real
// semantics in both builds.  CA_EMU keeps {pending arrivals, expected arrivals, phase} in the 64-bit word and spins by
// yielding to the other fibers of the block; a copy completes at issue there, so expect_tx is a plain arrival.
#if defined(CA_EMU)
static inline void ca_mbarx_init(unsigned long long* bar, int count) { *bar = ((unsigned long long)count << 32) | ((unsigned long long)count << 1); }
static inline void ca_mbarx_arrive(unsigned long long* bar) {
    unsigned long long v = *bar;
    unsigned pending = (unsigned)((v >> 1) & 0x7fffffffu), phase = (unsigned)(v & 1u), count = (unsigned)(v >> 32);
    if (--pending == 0) { pending = count; phase ^= 1u; }
    *bar = ((unsigned long long)count << 32) | ((unsigned long long)pending << 1) | phase;
}
static inline void ca_mbarx_expect_tx(unsigned long long* bar, unsigned bytes) { (void)bytes; ca_mbarx_arrive(bar); }
static inline void ca_mbarx_wait(unsigned long long* bar, unsigned parity) { while ((unsigned)(*bar & 1u) == (parity & 1u)) emu::yield(); }
#else
static __device__ __forceinline__ void ca_mbarx_init(unsigned long long* bar, int count) { ca_mbar_init(bar, count); }
static __device__ __forceinline__ void ca_mbarx_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ca_smem_u32(bar)) : "memory");
}
static __device__ __forceinline__ void ca_mbarx_expect_tx(unsigned long long* bar, unsigned bytes) { ca_mbar_expect_tx(bar, bytes); }
static __device__ __forceinline__ void ca_mbarx_wait(unsigned long long* bar, unsigned parity) { ca_mbar_wait(bar, parity); }
#endif

// ---- packed fp32 pairs (sm_100: add/sub/mul/fma .rn.f32x2): IEEE round-to-nearest per lane, i.e. bit-identical to the
// scalar _rn intrinsics, at one issue slot for two values (the kernels that use them are bound by issue slots)
#if defined(CA_EMU)
static inline float2 ca_sub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
static inline float2 ca_add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 ca_mul2(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 ca_fma2(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#else
static __device__ __forceinline__ float2 ca_sub2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
static __device__ __forceinline__ float2 ca_add2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
static __device__ __forceinline__ float2 ca_mul2(float2 a, float2 b) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
static __device__ __forceinline__ float2 ca_fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
#endif
