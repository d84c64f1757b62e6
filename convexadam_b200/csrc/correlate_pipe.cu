// correlate() as ONE launch of short-lived producer CTAs and marching consumer CTAs that hand the raw SSD planes over
// through a ring of a few planes in L2 -- the raw planes never reach HBM and nothing is recomputed.
//
// Replaces, for the certified mode (ca_correlate_fast), the pair k_raw_ssd_tmap -> 3 GB chunk buffer -> k_box12_fast of
// correlate.cu, whose raw planes made a round trip through HBM (27 GB of DRAM traffic per direction at 192x160x256 for
// 8.7 GB of algorithmic bytes).  Reference: `correlate`, src/convexAdam/convex_adam_utils.py:72-89.
//
// Mathematics (as correlate_fused.cu).  ssd = box3(box3(raw)) with zero padding and the intermediate truncated to the
// volume (utils.py:84-85) = T_H T_W T_D raw / 729 with the 1-D operator T = B M B (B: zero-padded 3-tap sum, M: mask to
// [0, n)).  Written out, (T r)(i) = r(i-2) + 2 r(i-1) + c(i) r(i) + 2 r(i+1) + r(i+2) on the zero-extended r, with the centre
// weight c(i) = 3 - [i = 0] - [i = n-1]: the truncation only removes the first-box value that sits outside the volume.
//
// Work decomposition.  The labels are processed in chunks of one H shift kH x NW (<= 4) W shifts x all S = 2*hw+1 D shifts,
// the H axis in blocks of PB planes.  Every CTA draws a ticket (one atomic) and the ticket decides what it is:
//   producer (plane x, 4 rows): the raw-SSD arithmetic of k_raw_ssd_tmap -- fixed tile and moving tile with its
//             displacement neighbourhood by tensor-map TMA from the unpadded features (out-of-bounds zero fill = F.pad,
//             utils.py:76), 4 voxels x S shifts per thread as packed pairs of neighbouring labels -- stored into the
//             ring slot of plane x once the consumers of the plane that held the slot before are done.
//   consumer (label, tile of 16 rows, block of PB planes): a loader thread waits until the producers have finished the
//             block's PB+4 planes and fetches the tile's 20 rows of each with ONE bulk asynchronous copy (cp.async.bulk +
//             mbarrier, CP_NST planes in flight); the compute threads apply T_D to the 20 rows (through a second
//             shared-memory plane), T_W for their 4 rows x 4 voxels, T_H along the block with the last four in-plane
//             results held in registers (the march along H, the axis the reference loops over, utils.py:83), and stream
//             the block's PB planes into `ssd`.  The 4 halo planes of a block are filtered in-plane twice (by two
//             blocks): (PB+4)/PB of the consumers' in-plane arithmetic, nothing of the raw SSD is recomputed.
// When a row of the volume is exactly one warp of quads (d = 128) the producers apply T_D themselves, with the two voxels
// either side of a thread's quad taken from the neighbouring lanes by warp shuffles; the consumers then skip their T_D phase
// and its shared-memory plane and keep six planes in flight instead of four (6.8 -> 6.5 ms per direction at config 2).
// Tickets are handed out, per chunk, in the order [producers of the planes block 0 needs][consumers of block 0]
// [producers of the further planes block 1 needs][consumers of block 1]...  With a ring of R >= PB + 4 planes a CTA only
// ever waits for CTAs with SMALLER tickets -- which hold an SM slot already or are done -- so there is no deadlock whatever
// order the hardware starts CTAs in, however many launches share the GPU, and every CTA is short-lived: the hardware
// scheduler balances the load.  (Measured and rejected: interleaving the consumers of block b-1 with the producers of block
// b ticket by ticket, so that both roles share every SM at any time: the producers then wait for ring slots twice as long,
// 10.0 instead of 8.4 ms per direction at the same ring depth.)
//
// Arithmetic: FMA channel sum, 5-tap sums, one multiplication by 1/729 -- at most C + 14 roundings per entry on sums of
// non-negative terms, i.e. within the same bound as the two-kernel fast path (coupled.cu, "certified selection").  The
// argmin / selections are never taken from this volume directly: see ca_coupled_convex_certified.
#include "ca_tma.cuh"

#include <cstdlib>
#ifdef CA_EMU
#include <sched.h>
#endif

#define CP_TY 4          // producer: rows per item
#define CP_TYC 16        // consumer: output rows per tile (4 row groups of 4 rows)
#define CP_TR 20         // consumer: ring rows per tile (2 + 16 + 2)
#define CP_NST 4         // consumer: planes in flight (its own T_D phase, two T_D planes next to the stages)
#define CP_NST_TDP 6     // ... when the producers filter along D (no T_D planes)
#define CP_NSTMAX 6
#define CP_RMAX 40       // ring depth (planes) the workspace is sized for
#define CP_MAXB 256      // H blocks per chunk
#define CP_MAXG 17       // kW groups (S <= 17)

struct CorrPipeArgs {
    float* ring;
    float* ssd;
    unsigned* ctr;            // [0] ticket, [1] unused, [2, 2+nseq) producer tiles done per plane, then consumers done per (chunk, block)
    int PB, nb;               // planes per H block, blocks
    int C, h, w, d;
    int xb, xe, r0, r1, nx;   // output planes [xb, xe); raw planes [r0, r1) = [xb-2, xe+2) clipped to the volume; nx = r1 - r0
    int nq, ncomp, ldr;       // quads per row, compute threads (4 nq), first thread of the loader warp
    int ntP, ntC;             // producer row tiles, consumer row tiles
    int R, pitch;             // ring depth; ring row pitch d + 8 (data at column 4)
    int labelPlane;           // (w + 4) * pitch floats: rows y = -2 .. w+1 of one label
    long long slotFloats;     // NWmax * S * labelPlane: one plane of a chunk
    int fix_f, mz, mrows;     // producer shared memory: offset of the moving tile, its row length and rows
    int fhalf, mhalf;         // ... floats between the two channel halves of the fixed / moving tile (128-byte multiples)
    int stage_f;              // consumer shared memory: floats per stage = CP_TR * pitch
    int nseq;                 // planes of all chunks: nchunks * nx
    int ng;                   // kW groups
    int nw[CP_MAXG], kw0[CP_MAXG];
    unsigned tstart[CP_MAXG + 1];   // first ticket of group g
    int ringRows;             // rows of one label in a ring slot: 16 ntC + 4 (every tile copies 20 whole rows)
    float* state;             // hand-over of the march state between the H blocks of a column (nullptr: blocks recompute their halo)
    int ncolsmax;             // columns (label x row tile) of the largest chunk
    unsigned* colflag;        // [chunk][column]: blocks of the column completed
    unsigned spin_limit;
};

// ---- hand-off between CTAs ----------------------------------------------------------------------------------------------
#if defined(CA_EMU)
static unsigned g_cp_err = 0;
static inline unsigned cp_ld_acquire(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void cp_backoff() { emu::yield(); sched_yield(); }
static inline void cp_fence_async() {}
static inline void cp_threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void cp_signal(unsigned* p) { __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL); }
static inline void cp_signal_relaxed(unsigned* p) { __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL); }
static inline unsigned cp_ld_relaxed(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
static inline void cp_fence_acquire() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned cp_ticket(unsigned* p) { return __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL); }
static inline void cp_st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
static inline float4 cp_ld4_cg(const float4* p) { return *p; }
static inline void cp_st4_cg(float4* p, const float4& v) { *p = v; }
#else
__device__ unsigned g_cp_err;
static __device__ __forceinline__ unsigned cp_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ void cp_backoff() { __nanosleep(64); }
// the bulk copy engine reads through the async proxy: order it after the acquire of the producers' generic-proxy stores
static __device__ __forceinline__ void cp_fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
static __device__ __forceinline__ void cp_threadfence() { __threadfence(); }
static __device__ __forceinline__ void cp_signal(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
// a ring slot is handed back once its copy has LANDED in shared memory (observed on the mbarrier): nothing of this thread
// needs to be ordered before the increment, so no release fence
static __device__ __forceinline__ void cp_signal_relaxed(unsigned* p) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory"); }
static __device__ __forceinline__ unsigned cp_ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
static __device__ __forceinline__ void cp_fence_acquire() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
static __device__ __forceinline__ unsigned cp_ticket(unsigned* p) { return atomicAdd(p, 1u); }
// the march state handed from one H block of a column to the next: written by one CTA, read by another -> L2 (.cg), not L1
static __device__ __forceinline__ float4 cp_ld4_cg(const float4* p) { return __ldcg(p); }
static __device__ __forceinline__ void cp_st4_cg(float4* p, const float4& v) { __stcg(p, v); }
// ring stores: plain write-back stores (the lines are meant to stay in L2 until the consumers have fetched them)
static __device__ __forceinline__ void cp_st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
#endif

#ifdef CP_PROFILE
// debug build only (tools/build_cp_profile.sh): cycles per phase of one thread per role, summed over all CTAs
__device__ unsigned long long cp_prof[24];
#define CPP_T(var) const long long var = clock64()
#define CPP_ACC(slot, t0, t1) atomicAdd(&cp_prof[slot], (unsigned long long)((t1) - (t0)))
#define CPP_CNT(slot) atomicAdd(&cp_prof[slot], 1ull)
extern "C" __attribute__((visibility("default"))) int ca_debug_cp_prof(unsigned long long* out, int reset) {
    unsigned long long z[24] = {0};
    if (cudaMemcpyFromSymbol(out, cp_prof, sizeof(z)) != cudaSuccess) return -1;
    if (reset) cudaMemcpyToSymbol(cp_prof, z, sizeof(z));
    return 0;
}
#else
#define CPP_T(var)
#define CPP_ACC(slot, t0, t1)
#define CPP_CNT(slot)
#endif

// wait until *p >= target.  A wait that outlasts spin_limit polls (seconds) sets the error flag and gives up: the results
// are then garbage, but the kernel ends (tests read the flag through ca_debug_corr_pipe_error).
static __device__ __forceinline__ void cp_poll(const unsigned* p, unsigned target, unsigned limit) {
    unsigned* err = &g_cp_err;
    unsigned spins = 0;
    while (cp_ld_acquire(p) < target) {
        cp_backoff();
        if (++spins > limit || ((spins & 1023u) == 0 && cp_ld_acquire(err) != 0)) {
#if defined(CA_EMU)
            __atomic_store_n(err, 1u, __ATOMIC_RELEASE);
#else
            atomicExch(err, 1u);
#endif
            return;
        }
    }
}

// how many of the counters p[0 .. n) have reached target, counted from p[0] (n <= 8); waits until at least one has.
// One round trip to L2 for the whole batch, one acquire fence for all of them.
static __device__ __forceinline__ int cp_poll_prefix(const unsigned* p, int n, unsigned target, unsigned limit) {
    unsigned* err = &g_cp_err;
    unsigned spins = 0;
    for (;;) {
        unsigned v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = k < n ? cp_ld_relaxed(p + k) : 0u;
        int ok = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) ok += (ok == k && k < n && v[k] >= target) ? 1 : 0;
        if (ok > 0) { cp_fence_acquire(); return ok; }
        cp_backoff();
        if (++spins > limit || ((spins & 1023u) == 0 && cp_ld_acquire(err) != 0)) {
#if defined(CA_EMU)
            __atomic_store_n(err, 1u, __ATOMIC_RELEASE);
#else
            atomicExch(err, 1u);
#endif
            return n;
        }
    }
}

// ---- producer: raw SSD of 4 rows of plane x for one kH, NWc values of kW and all S values of kD -------------------------
template <int HW, bool TDP>
static __device__ __forceinline__ void cp_producer(const CaTensorMap* tmF, const CaTensorMap* tmM, const CorrPipeArgs& a, float* sm,
                                                   unsigned long long* bars, int tid, int chunk, int g, int kH, int item) {
    constexpr int S = 2 * HW + 1;
    constexpr int MSH = (4 - HW % 4) % 4;
    constexpr int NM = (MSH + 4 + 2 * HW + 3) & ~3;
    const int NWc = a.nw[g], kW0 = a.kw0[g];
    const int xi = item / a.ntP, tile = item - xi * a.ntP;
    const int x = a.r0 + xi, y0t = tile * CP_TY;
    const unsigned seq = (unsigned)chunk * (unsigned)a.nx + (unsigned)xi;
    float* sfix = sm;                                                // [C][TY][d]
    float* smov = sm + a.fix_f;                                      // [C][mrows][mz]
    const int fch = CP_TY * a.d, mch = a.mrows * a.mz;
    // the tiles arrive in two halves of the channels (two tensor-map boxes each): the arithmetic starts on the first half
    // while the second is still in flight
    unsigned long long* tile_bar = &bars[0];
    unsigned long long* slot_bar = &bars[1];
    unsigned long long* tile_bar2 = &bars[2];
    const int Ch = (a.C + 1) / 2;
    if (tid == 0) { ca_mbarx_init(tile_bar, 1); ca_mbarx_init(slot_bar, 1); ca_mbarx_init(tile_bar2, 1); }
    __syncthreads();
    CPP_T(p0);
    if (tid == a.ldr) {
        ca_mbarx_expect_tx(tile_bar, (unsigned)(Ch * (fch + mch)) * 4u);
        ca_tma_load_4d(sfix, tmF, 0, y0t, x, 0, tile_bar);
        ca_tma_load_4d(smov, tmM, -HW - MSH, y0t + kW0 - HW, x + kH - HW, 0, tile_bar);
        ca_mbarx_expect_tx(tile_bar2, (unsigned)(Ch * (fch + mch)) * 4u);
        ca_tma_load_4d(sfix + a.fhalf, tmF, 0, y0t, x, Ch, tile_bar2);
        ca_tma_load_4d(smov + a.mhalf, tmM, -HW - MSH, y0t + kW0 - HW, x + kH - HW, Ch, tile_bar2);
        // the ring slot of this plane held plane seq-R: the (at most two) consumer blocks that read THAT plane must be done
        if (seq >= (unsigned)a.R) {
            const unsigned old = seq - (unsigned)a.R;
            const int oc = (int)(old / (unsigned)a.nx);                       // its chunk
            const int q = a.r0 + (int)(old - (unsigned)oc * (unsigned)a.nx);   // its plane
            const unsigned target = (unsigned)(a.nw[oc / S] * S * a.ntC);
            const unsigned* cd = a.ctr + 2 + a.nseq + oc * a.nb;
            if (a.state) {
                // with the hand-over a plane is read by one block only: block b fetches the planes [ob+2, oe+2) (block 0 also the first)
                const int b = ca_clamp((q - a.xb - 2 + a.PB) / a.PB - 1, 0, a.nb - 1);
                cp_poll(cd + b, target, a.spin_limit);
            } else {
                const int b_last = ca_min(a.nb - 1, (q - a.xb + 2) / a.PB);
                if (b_last > 0) cp_poll(cd + b_last - 1, target, a.spin_limit);
                cp_poll(cd + b_last, target, a.spin_limit);
            }
        }
        CPP_T(p1);
        CPP_ACC(6, p0, p1);
        ca_mbarx_arrive(slot_bar);
    }
    if (tid < a.ncomp) {
        const int tx = tid % a.nq, ty = tid / a.nq;
        const int z0 = 4 * tx, y = y0t + ty;
        const bool active = y < a.w;
        ca_mbarx_wait(tile_bar, 0);
        CPP_T(p2);
        if (tid == 0) CPP_ACC(0, p0, p2);
        const float* fs0 = sfix + ty * a.d + z0;
        float* rp0 = a.ring + (long long)(seq % (unsigned)a.R) * a.slotFloats + (long long)(y + 2) * a.pitch + 4 + z0;
        for (int kWl = 0; kWl < NWc; ++kWl) {
            const float* ms0 = smov + (ty + kWl) * a.mz + z0;
            // packed pairs of neighbouring labels of one voxel (see k_raw_ssd_staged): first moving index of voxel v, label k
            // is MSH + v + k
            float2 pr[4][S / 2];
            float sg[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                sg[v] = 0.f;
#pragma unroll
                for (int q = 0; q < S / 2; ++q) pr[v][q] = make_float2(0.f, 0.f);
            }
            auto accumulate = [&](const float* fs, const float* ms, int nch) {
#pragma unroll 2
                for (int c = 0; c < nch; ++c) {
                    const float4 f4 = *reinterpret_cast<const float4*>(fs + c * fch);
                    const float f[4] = {f4.x, f4.y, f4.z, f4.w};
                    float m[NM];
#pragma unroll
                    for (int j = 0; j < NM / 4; ++j) {
                        const float4 m4 = *reinterpret_cast<const float4*>(ms + c * mch + 4 * j);
                        m[4 * j + 0] = m4.x; m[4 * j + 1] = m4.y; m[4 * j + 2] = m4.z; m[4 * j + 3] = m4.w;
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float2 fv = make_float2(f[v], f[v]);
                        const int par = (MSH + v) & 1;
#pragma unroll
                        for (int q = 0; q < S / 2; ++q) {
                            const int i0 = MSH + v + 2 * q + par;
                            const float2 df = ca_sub2(make_float2(m[i0], m[i0 + 1]), fv);
                            pr[v][q] = ca_fma2(df, df, pr[v][q]);
                        }
                        const float ds = m[MSH + v + (par ? 0 : S - 1)] - f[v];
                        sg[v] = __fmaf_rn(ds, ds, sg[v]);
                    }
                }
            };
            accumulate(fs0, ms0, Ch);
            if (kWl == 0) ca_mbarx_wait(tile_bar2, 0);
            accumulate(fs0 + a.fhalf, ms0 + a.mhalf, a.C - Ch);
            if (kWl == 0) {
                CPP_T(p3);
                ca_mbarx_wait(slot_bar, 0);
                CPP_T(p4);
                if (tid == 0) CPP_ACC(2, p3, p4);
            }
            {
                float* rp = rp0 + (long long)(kWl * S) * a.labelPlane;
#pragma unroll
                for (int k = 0; k < S; ++k) {
                    float o[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        if (((MSH + v) & 1) == 0) o[v] = k == S - 1 ? sg[v] : ((k & 1) ? pr[v][k >> 1].y : pr[v][k >> 1].x);
                        else o[v] = k == 0 ? sg[v] : (((k - 1) & 1) ? pr[v][(k - 1) >> 1].y : pr[v][(k - 1) >> 1].x);
                    }
                    if (TDP) {
                        // T_D here: a row is exactly one warp (32 quads), so the two voxels either side of the thread's quad come
                        // from the neighbouring lanes (zero beyond the row ends = the zero padding); same operations in the same
                        // order as cp_td, which the consumers then skip
                        float l2 = __shfl_up_sync(0xffffffffu, o[2], 1), l3 = __shfl_up_sync(0xffffffffu, o[3], 1);
                        float r0 = __shfl_down_sync(0xffffffffu, o[0], 1), r1 = __shfl_down_sync(0xffffffffu, o[1], 1);
                        if (tx == 0) { l2 = 0.f; l3 = 0.f; }
                        if (tx == a.nq - 1) { r0 = 0.f; r1 = 0.f; }
                        const float2 two = make_float2(2.f, 2.f);
                        const float2 ae01 = ca_add2(make_float2(l2, l3), make_float2(o[2], o[3]));
                        const float2 ae23 = ca_add2(make_float2(o[0], o[1]), make_float2(r0, r1));
                        const float b0 = l3 + o[1], b1 = o[0] + o[2], b2 = o[1] + o[3], b3 = o[2] + r0;
                        const float2 d01 = ca_fma2(make_float2(tx == 0 ? 2.f : 3.f, 3.f), make_float2(o[0], o[1]), ca_fma2(two, make_float2(b0, b1), ae01));
                        const float2 d23 = ca_fma2(make_float2(3.f, tx == a.nq - 1 ? 2.f : 3.f), make_float2(o[2], o[3]), ca_fma2(two, make_float2(b2, b3), ae23));
                        o[0] = d01.x; o[1] = d01.y; o[2] = d23.x; o[3] = d23.y;
                    }
                    if (active) cp_st4(rp + (long long)k * a.labelPlane, make_float4(o[0], o[1], o[2], o[3]));
                }
            }
        }
        CPP_T(p5);
        if (tid == 0) CPP_ACC(1, p2, p5);
    }
    // the CTA barrier orders every thread's ring stores before thread 0's release-increment (cumulative): one release per
    // CTA, no fence per thread (a gpu-scope fence also invalidates L1)
    __syncthreads();
    if (tid == 0) {
        CPP_T(p5b);
        cp_signal(a.ctr + 2 + seq);
        CPP_T(p6);
        CPP_ACC(3, p5b, p6);
        CPP_T(p7);
        CPP_ACC(5, p0, p7);
        CPP_CNT(4);
    }
}

// ---- consumer: T_W T_D inside a plane, T_H across planes, for one label and 16 rows --------------------------------------
// phase A: T_D of one ring row: 8 voxels (z0-2 .. z0+5) in, the 4 filtered values of z0 .. z0+3 out
static __device__ __forceinline__ float4 cp_td(const float* r, const float (&wz)[4]) {
    const float2 two = make_float2(2.f, 2.f);
    const float2 lo = *reinterpret_cast<const float2*>(r - 2);
    const float4 mid = *reinterpret_cast<const float4*>(r);
    const float2 hi = *reinterpret_cast<const float2*>(r + 4);
    const float2 ae01 = ca_add2(lo, make_float2(mid.z, mid.w));
    const float2 ae23 = ca_add2(make_float2(mid.x, mid.y), hi);
    const float b0 = lo.y + mid.y, b1 = mid.x + mid.z, b2 = mid.y + mid.w, b3 = mid.z + hi.x;
    const float2 d01 = ca_fma2(make_float2(wz[0], wz[1]), make_float2(mid.x, mid.y), ca_fma2(two, make_float2(b0, b1), ae01));
    const float2 d23 = ca_fma2(make_float2(wz[2], wz[3]), make_float2(mid.z, mid.w), ca_fma2(two, make_float2(b2, b3), ae23));
    return make_float4(d01.x, d01.y, d23.x, d23.y);
}
// barrier among the compute threads of a consumer (the loader thread is busy elsewhere): a named hardware barrier when they
// are whole warps, else an mbarrier with one arrival per warp
static __device__ __forceinline__ void cp_compute_barrier(unsigned long long* dbar, unsigned parity, int ncomp, unsigned cmask, bool elect) {
#if !defined(CA_EMU)
    if ((ncomp & 31) == 0) {
        asm volatile("bar.sync 1, %0;" ::"r"(ncomp) : "memory");
        return;
    }
#endif
    __syncwarp(cmask);
    if (elect) ca_mbarx_arrive(dbar);
    ca_mbarx_wait(dbar, parity);
}

struct CpConsumerThread {
    const float* stageA;      // phase A: the thread's first ring row (tile row 5*rg) at column z0 in stage 0
    float* dA;                // phase A: where its T_D rows go (buffer 0)
    const float* dB;          // phase B: the thread's first T_D row (tile row 4*rg) at column z0 (buffer 0)
    float* out;               // the thread's first row of the cost volume at the next plane to store
    float wz0, wz3;           // centre weights of T_D at the thread's first / last voxel (the two in between: 3)
    int yb;                   // the thread's first output row
    int nrows;                // its rows inside the volume (0..4)
    int64_t outPitch;
    unsigned cmask;           // compute lanes of the thread's warp
    bool elect;               // first lane of its warp: arrives on the barriers for the warp
    bool prof;                // CP_PROFILE: the thread that reports
    int rs, re, ob, oe;       // the item's block: raw planes [rs, re) in, cost-volume planes [ob, oe) out
};

// Phase A (T_D of the ring rows into the T_D buffer) runs one plane AHEAD of phase B (T_W out of the T_D buffer, T_H,
// store): step X does A(X+1) and B(X) between the same pair of barriers, so the two phases overlap in the instruction stream.
static __device__ __forceinline__ void cp_phase_a(int X, const CorrPipeArgs& a, const CpConsumerThread& t, unsigned long long* full,
                                                  unsigned long long* empty) {
    if (X >= t.re) return;
    const int i = X - t.rs, st = i % CP_NST;
    const int db = (i & 1) * CP_TR * a.d;
    CPP_T(c0);
    ca_mbarx_wait(&full[st], (unsigned)(i / CP_NST) & 1u);
    CPP_T(c1);
    const float* sp = t.stageA + st * a.stage_f;
    const float wz[4] = {t.wz0, 3.f, 3.f, t.wz3};
#pragma unroll
    for (int jj = 0; jj < 5; ++jj) *reinterpret_cast<float4*>(t.dA + db + jj * a.d) = cp_td(sp + jj * a.pitch, wz);
    __syncwarp(t.cmask);
    if (t.elect) ca_mbarx_arrive(&empty[st]);
    CPP_T(c2);
    if (t.prof) { CPP_ACC(8, c0, c1); CPP_ACC(9, c1, c2); CPP_CNT(12); }
}

// one plane step; PH = (X - rs) mod 4 is the register slot that holds the in-plane result of plane X-4
template <int PH, bool TDP>
static __device__ __forceinline__ void cp_consumer_step(int X, const CorrPipeArgs& a, CpConsumerThread& t, unsigned long long* full,
                                                        unsigned long long* empty, unsigned long long* dbar, float2 (&hs)[4][4][2]) {
    if (X > t.oe + 1) return;
    const float2 two = make_float2(2.f, 2.f);
    const bool have = X < t.re;
    const int i = X - t.rs;
    // TDP: the ring holds T_D values already (the producers filtered along D): phase B reads the stage itself
    constexpr int NST = TDP ? CP_NST_TDP : CP_NST;
    const int db = TDP ? (i % NST) * a.stage_f : (i & 1) * CP_TR * a.d;
    const int dpitch = TDP ? a.pitch : a.d;
    if (TDP) {
        if (have) {
            CPP_T(c0);
            ca_mbarx_wait(&full[i % NST], (unsigned)(i / NST) & 1u);
            CPP_T(c1);
            if (t.prof) { CPP_ACC(8, c0, c1); CPP_CNT(12); }
        }
    } else {
        cp_phase_a(X + 1, a, t, full, empty);
    }
    CPP_T(c4);
    const int xo = X - 2;
    const bool store = xo >= t.ob;
    const float wx = 3.f - (xo == 0 ? 1.f : 0.f) - (xo == a.h - 1 ? 1.f : 0.f);
    const float c729 = 1.0f / 729.0f;
    const float2 wx2 = make_float2(wx, wx), c2 = make_float2(c729, c729);
    // phase B: T_W (8 rows of T_D values in, tile rows 4*rg .. 4*rg+7; row o of the thread is complete after row o+4),
    // then T_H of the finished row straight away
    float2 P[4][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float2 D[2];
        if (have) {
            const float4 v = *reinterpret_cast<const float4*>(t.dB + db + j * dpitch);
            D[0] = make_float2(v.x, v.y); D[1] = make_float2(v.z, v.w);
        } else {
            D[0] = D[1] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int k = j - o;
            if (k < 0 || k > 4) continue;
            const int y = t.yb + o;
            const float wy = 3.f - (y == 0 ? 1.f : 0.f) - (y == a.w - 1 ? 1.f : 0.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (k == 0) P[o][h] = D[h];
                else if (k == 4) P[o][h] = ca_add2(P[o][h], D[h]);
                else if (k == 2) P[o][h] = ca_fma2(make_float2(wy, wy), D[h], P[o][h]);
                else P[o][h] = ca_fma2(two, D[h], P[o][h]);
            }
        }
        if (j >= 4) {
            const int o = j - 4;
            if (store) {
                float2 r[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 ae = ca_add2(hs[PH][o][h], P[o][h]);
                    const float2 bd = ca_add2(hs[(PH + 1) & 3][o][h], hs[(PH + 3) & 3][o][h]);
                    r[h] = ca_mul2(ca_fma2(wx2, hs[(PH + 2) & 3][o][h], ca_fma2(two, bd, ae)), c2);
                }
                if (o < t.nrows) ca_st_stream4(reinterpret_cast<float4*>(t.out + o * a.d), make_float4(r[0].x, r[0].y, r[1].x, r[1].y));
            }
            hs[PH][o][0] = P[o][0]; hs[PH][o][1] = P[o][1];
        }
    }
    if (store) t.out += t.outPitch;
    CPP_T(c5);
    if (TDP) {
        if (have) {     // the stage has been read
            __syncwarp(t.cmask);
            if (t.elect) ca_mbarx_arrive(&empty[i % NST]);
        }
    } else {
        // A(X+1) is complete and B(X) has read its T_D plane: the next step may overwrite the latter and read the former
        if (X + 1 < t.re) cp_compute_barrier(dbar, (unsigned)(i + 1) & 1u, a.ncomp, t.cmask, t.elect);
    }
    CPP_T(c6);
    if (t.prof) { CPP_ACC(11, c4, c5); CPP_ACC(10, c5, c6); }
}

template <int HW, bool TDP>
static __device__ __forceinline__ void cp_consumer(const CorrPipeArgs& a, float* sm, unsigned long long* bars, int tid, int chunk, int g,
                                                   int kH, int blk, int col) {
    constexpr int S = 2 * HW + 1;
    const int j = col / a.ntC, tile = col - j * a.ntC;              // label of the chunk, row tile
    const int kWl = j / S, kD = j - kWl * S;
    const int64_t label = (int64_t)kD * S * S + (int64_t)(a.kw0[g] + kWl) * S + kH;
    const int ty0 = tile * CP_TYC;
    // the item's block of H planes
    const int ob = a.xb + blk * a.PB, oe = ca_min(ob + a.PB, a.xe);
    // with the hand-over, a block other than the first starts from its predecessor's last four in-plane results (the planes
    // ob-2 .. ob+1) and filters the planes from ob+2 on; else it filters its own 2 + 2 halo planes again
    const bool hand = TDP && a.state != nullptr;
    const bool from_state = hand && blk > 0;
    const int rs = from_state ? ob + 2 : ca_max(a.r0, ob - 2), re = ca_min(a.r1, oe + 2);
    unsigned long long* full = bars;
    constexpr int NST = TDP ? CP_NST_TDP : CP_NST;
    unsigned long long* empty = bars + CP_NSTMAX;
    unsigned long long* dbar = bars + 2 * CP_NSTMAX;
    if (tid == 0) {
        const int ncw = (a.ncomp + 31) / 32;             // compute warps: one arrival each
        for (int s = 0; s < NST; ++s) { ca_mbarx_init(&full[s], 1); ca_mbarx_init(&empty[s], ncw); }
        ca_mbarx_init(dbar, ncw);
    }
    __syncthreads();
    const int nraw = ca_max(re - rs, 0);
    CPP_T(q0);
    if (tid == a.ldr) {
        // ring rows [ty0, ty0+20) of the label <-> y = ty0-2 .. ty0+17; rows 0, 1 and >= w+2 and the pad columns are zero
        const unsigned bytes = (unsigned)(CP_TR * a.pitch) * 4u;
        const long long off = (long long)j * a.labelPlane + (long long)ty0 * a.pitch;
        const unsigned seq0 = (unsigned)chunk * (unsigned)a.nx + (unsigned)(rs - a.r0);
        int ready = 0;                                  // planes [0, ready) of the block are known to be complete in the ring
        for (int i = 0; i < nraw; ++i) {
            const int st = i % NST;
            if (i >= NST) ca_mbarx_wait(&empty[st], (unsigned)(i / NST - 1) & 1u);
            const unsigned seq = seq0 + (unsigned)i;
            if (i >= ready) {
                ready = i + cp_poll_prefix(a.ctr + 2 + seq, ca_min(8, nraw - i), (unsigned)a.ntP, a.spin_limit);
                cp_fence_async();
            }
            ca_mbarx_expect_tx(&full[st], bytes);
            ca_bulk_g2s(sm + st * a.stage_f, a.ring + (long long)(seq % (unsigned)a.R) * a.slotFloats + off, bytes, &full[st]);
        }
    }
    if (tid < a.ncomp) {
        const int q = tid % a.nq, rg = tid / a.nq;
        const int z0 = 4 * q;
        CpConsumerThread t;
        float* dbuf = sm + NST * a.stage_f;                          // [2][CP_TR][d]: the T_D rows of the plane in work
        t.stageA = sm + (5 * rg) * a.pitch + 4 + z0;
        t.dA = dbuf + (5 * rg) * a.d + z0;
        t.dB = TDP ? sm + (4 * rg) * a.pitch + 4 + z0 : dbuf + (4 * rg) * a.d + z0;
        t.outPitch = (int64_t)a.w * a.d;
        t.prof = tid == 0;
        t.elect = (tid & 31) == 0;
        t.cmask = (tid | 31) < a.ncomp ? 0xffffffffu : ((1u << (a.ncomp & 31)) - 1u);
        t.wz0 = z0 == 0 ? 2.f : 3.f;
        t.wz3 = z0 + 3 == a.d - 1 ? 2.f : 3.f;
        t.yb = ty0 + 4 * rg;
        t.nrows = ca_clamp(a.w - t.yb, 0, 4);
        t.rs = rs; t.re = re; t.ob = ob; t.oe = oe;
        t.out = a.ssd + ((label * (a.xe - a.xb) + (ob - a.xb)) * a.w + (t.nrows > 0 ? t.yb : 0)) * (int64_t)a.d + z0;
        float2 hs[4][4][2];
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int o = 0; o < 4; ++o) hs[s][o][0] = hs[s][o][1] = make_float2(0.f, 0.f);
        unsigned nbar = 0;                      // barriers of the compute threads so far (TDP; the other path counts planes)
        unsigned* flag = a.colflag + (size_t)chunk * a.ncolsmax + col;
        if (!TDP) {
            cp_phase_a(rs, a, t, full, empty);
            cp_compute_barrier(dbar, 0u, a.ncomp, t.cmask, t.elect);
        }
        if (from_state) {
            // the predecessor (same column, block blk-1: a smaller ticket) has stored its state and released the flag
            if (tid == 0) cp_poll(flag, (unsigned)blk, a.spin_limit);
            cp_compute_barrier(dbar, nbar++ & 1u, a.ncomp, t.cmask, t.elect);
            const float4* sp = reinterpret_cast<const float4*>(a.state) +
                               ((size_t)((chunk & 1) * 2 + ((blk - 1) & 1)) * a.ncolsmax + col) * (size_t)(16 * a.ncomp) + tid;
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const float4 v = cp_ld4_cg(sp + (k * 4 + o) * a.ncomp);
                    hs[k][o][0] = make_float2(v.x, v.y); hs[k][o][1] = make_float2(v.z, v.w);
                }
        }
        for (int X = rs; X <= oe + 1; X += 4) {
            cp_consumer_step<0, TDP>(X, a, t, full, empty, dbar, hs);
            cp_consumer_step<1, TDP>(X + 1, a, t, full, empty, dbar, hs);
            cp_consumer_step<2, TDP>(X + 2, a, t, full, empty, dbar, hs);
            cp_consumer_step<3, TDP>(X + 3, a, t, full, empty, dbar, hs);
        }
        CPP_T(q1);
        if (tid == 0) { CPP_ACC(13, q0, q1); CPP_CNT(14); }
        if (hand) {
            if (blk < a.nb - 1) {
                // the state buffer of (chunk parity, block parity) was last read by block blk+1 of chunk-2: that one must be done
                if (tid == 0 && chunk >= 2 && col < a.nw[(chunk - 2) / S] * S * a.ntC)
                    cp_poll(a.colflag + (size_t)(chunk - 2) * a.ncolsmax + col, (unsigned)ca_min(blk + 2, a.nb), a.spin_limit);
                cp_compute_barrier(dbar, nbar++ & 1u, a.ncomp, t.cmask, t.elect);
                float4* sp = reinterpret_cast<float4*>(a.state) +
                             ((size_t)((chunk & 1) * 2 + (blk & 1)) * a.ncolsmax + col) * (size_t)(16 * a.ncomp) + tid;
                // oldest first: after n = oe + 2 - rs steps the in-plane result of plane oe-2+k sits in slot (n + k) mod 4
                const int rot = (oe + 2 - rs) & 3;
#define CP_STORE_STATE(R)                                                                                              \
    _Pragma("unroll") for (int k = 0; k < 4; ++k) _Pragma("unroll") for (int o = 0; o < 4; ++o)                        \
        cp_st4_cg(sp + (k * 4 + o) * a.ncomp, make_float4(hs[(R + k) & 3][o][0].x, hs[(R + k) & 3][o][0].y,            \
                                                          hs[(R + k) & 3][o][1].x, hs[(R + k) & 3][o][1].y))
                if (rot == 0) { CP_STORE_STATE(0); } else if (rot == 1) { CP_STORE_STATE(1); } else if (rot == 2) { CP_STORE_STATE(2); } else { CP_STORE_STATE(3); }
#undef CP_STORE_STATE
            }
            // every plane of the block has been fetched and used, the state is stored: release the column's flag (the barrier
            // orders every thread's stores before thread 0's release) and hand the ring planes back
            cp_compute_barrier(dbar, nbar++ & 1u, a.ncomp, t.cmask, t.elect);
            if (tid == 0) {
                cp_signal(flag);
                cp_signal_relaxed(a.ctr + 2 + a.nseq + chunk * a.nb + blk);
            }
        } else {
            // every plane of the block has been fetched and used by every compute thread: the item is done with the ring
            // (barrier number nraw of the compute threads: the prologue's and one per step that had a plane to prepare)
            cp_compute_barrier(dbar, TDP ? 0u : (unsigned)nraw & 1u, a.ncomp, t.cmask, t.elect);
            if (tid == 0) cp_signal_relaxed(a.ctr + 2 + a.nseq + chunk * a.nb + blk);
        }
    }
}

#define CP_MAXT 288
#if defined(CA_EMU)
#define CP_REGCAP
#elif !defined(CP_REGCAP)
#define CP_REGCAP __maxnreg__(128)      // 3 CTAs of 160 threads per SM at d = 128
#endif
template <int HW, bool TDP>
__global__ void CP_REGCAP k_corr_pipe(CA_TMAP_PARAM tmF, CA_TMAP_PARAM tmM, CorrPipeArgs a) {
    constexpr int S = 2 * HW + 1;
    CA_DYN_SMEM_ALIGNED(float, sm, 128);
    __shared__ unsigned long long bars[2 * CP_NSTMAX + 1];
    __shared__ int s_job[6];      // chunk, kW group, kH, role (0 producer / 1 consumer / -1 none), block, item
    const int tid = (int)threadIdx.x;
    if (tid == 0) {
        // thread 0 draws the ticket and decodes it for the CTA (divisions and a loop over the H blocks: not 160 times)
        const unsigned t = cp_ticket(a.ctr);
        int g = 0;
        while (g + 1 < a.ng && t >= a.tstart[g + 1]) ++g;
        const unsigned ncols = (unsigned)(a.nw[g] * S * a.ntC);
        const unsigned per_chunk = (unsigned)(a.nx * a.ntP) + (unsigned)a.nb * ncols;
        const unsigned u = t - a.tstart[g];
        const int kH = (int)(u / per_chunk);
        unsigned r = u - (unsigned)kH * per_chunk;
        s_job[0] = g * S + kH; s_job[1] = g; s_job[2] = kH; s_job[3] = -1;
        // within the chunk: [producers of the planes block 0 needs][consumers of block 0][producers of the further planes
        // block 1 needs][consumers of block 1] ...
        int pe_prev = a.r0;
        for (int b = 0; b < a.nb; ++b) {
            const int pe = b == a.nb - 1 ? a.r1 : ca_min(a.r1, a.xb + (b + 1) * a.PB + 2);
            const unsigned np = (unsigned)((pe - pe_prev) * a.ntP);
            if (r < np) { s_job[3] = 0; s_job[4] = b; s_job[5] = (pe_prev - a.r0) * a.ntP + (int)r; break; }
            r -= np;
            if (r < ncols) { s_job[3] = 1; s_job[4] = b; s_job[5] = (int)r; break; }
            r -= ncols;
            pe_prev = pe;
        }
    }
    __syncthreads();
    const int chunk = s_job[0], g = s_job[1], kH = s_job[2], role = s_job[3], blk = s_job[4], item = s_job[5];
#ifndef CP_NO_PRODUCER
    if (role == 0) cp_producer<HW, TDP>(&tmF, &tmM, a, sm, bars, tid, chunk, g, kH, item);
#endif
#ifndef CP_NO_CONSUMER
    if (role == 1) cp_consumer<HW, TDP>(a, sm, bars, tid, chunk, g, kH, blk, item);
#endif
}

// ---- host side ---------------------------------------------------------------------------------------------------------
namespace {
// CA_CORR_PIPE=1 switches the certified mode to this kernel (read on every call so that tests can toggle it).  Off by
// default: at 192x160x256 it measures 6.4 ms per direction against 6.3 ms for the two-kernel path, with 1.6-2.5x instead of
// 3.1x the algorithmic DRAM traffic (DESIGN.md 4.1 has the numbers and the reason).
bool pipe_enabled() {
    const char* e = getenv("CA_CORR_PIPE");
    return e && atoi(e) == 1;
}
int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
size_t cp_align256(size_t v) { return (v + 255) & ~(size_t)255; }
int g_last_cfg[6] = {0, 0, 0, 0, 0, 0};     // NW, PB, R, T_D in the producers, state hand-over, launches: what the last launch ran with

struct PipeGeom {
    int nq, ncomp, ldr, nt;
    int ntP, ntC, pitch, ringRows, labelPlane;
    int fix_f, fhalf, mz;
    size_t smem_cons, smem_cons_tdp;
};
bool pipe_geom(int C, int w, int d, int hw, PipeGeom* g) {
    if ((d & 3) || d > 240 || d < 4 || C > 24 || hw < 1 || hw > 8) return false;
    g->nq = d / 4; g->ncomp = 4 * g->nq; g->ldr = (g->ncomp + 31) & ~31; g->nt = g->ldr + 32;
    if (g->nt > CP_MAXT) return false;
    g->ntP = (w + CP_TY - 1) / CP_TY; g->ntC = (w + CP_TYC - 1) / CP_TYC;
    g->pitch = d + 8;
    g->ringRows = CP_TYC * g->ntC + 4;
    if ((int64_t)g->ringRows * g->pitch >= (1ll << 28)) return false;
    g->labelPlane = g->ringRows * g->pitch;
    const int msh = (4 - hw % 4) % 4, nm = (msh + 4 + 2 * hw + 3) & ~3;
    g->mz = d - 4 + nm;
    if (g->mz > 256) return false;
    g->fhalf = (((C + 1) / 2) * CP_TY * d + 31) & ~31;           // the tiles arrive as two boxes of ceil(C/2) channels
    g->fix_f = 2 * g->fhalf;
    g->smem_cons = ((size_t)CP_NST * CP_TR * g->pitch + (size_t)2 * CP_TR * d) * sizeof(float);
    g->smem_cons_tdp = (size_t)CP_NST_TDP * CP_TR * g->pitch * sizeof(float);
    return true;
}
size_t pipe_smem(const PipeGeom& g, int C, int NW, bool tdp) {
    const size_t mhalf = (size_t)((((C + 1) / 2) * (CP_TY + NW - 1) * g.mz + 31) & ~31);
    const size_t prod = ((size_t)g.fix_f + 2 * mhalf) * sizeof(float);
    const size_t cons = tdp ? g.smem_cons_tdp : g.smem_cons;
    return (prod > cons ? prod : cons) + 128;
}

template <int HW>
int launch_pipe(const float* fix, const float* mov, int C, int h, int w, int d, int xb, int xe, float* ssd, void* ws, size_t ws_bytes,
                cudaStream_t st) {
    constexpr int S = 2 * HW + 1;
    PipeGeom pg;
    if (!pipe_geom(C, w, d, HW, &pg)) return CA_ERR_UNSUPPORTED;
    // T_D inside the producers (warp shuffles) when a row of the volume is exactly one warp of quads; CA_PIPE_TDP=0: A/B switch
    const bool tdp = pg.nq == 32 && env_int("CA_PIPE_TDP", 1) != 0;
    auto k = tdp ? k_corr_pipe<HW, true> : k_corr_pipe<HW, false>;
    // kW values per chunk: the moving tile of a producer holds CP_TY + NW - 1 rows
    int NW = ca_min(ca_min(ca_max(env_int("CA_PIPE_NW", 3), 1), 4), S);
    while (NW > 1 && pipe_smem(pg, C, NW, tdp) > 75 * 1024) --NW;            // three CTAs per SM
    const size_t smem = pipe_smem(pg, C, NW, tdp);
    if (smem > 110 * 1024) return CA_ERR_UNSUPPORTED;
    CorrPipeArgs a;
    a.ssd = ssd; a.C = C; a.h = h; a.w = w; a.d = d;
    a.xb = xb; a.xe = xe;
    a.r0 = ca_max(0, xb - 2); a.r1 = ca_min(h, xe + 2); a.nx = a.r1 - a.r0;
    a.nq = pg.nq; a.ncomp = pg.ncomp; a.ldr = pg.ldr;
    a.ntP = pg.ntP; a.ntC = pg.ntC; a.pitch = pg.pitch; a.ringRows = pg.ringRows; a.labelPlane = pg.labelPlane;
    a.fix_f = pg.fix_f; a.mz = pg.mz; a.mrows = CP_TY + NW - 1;
    a.fhalf = pg.fhalf; a.mhalf = (((C + 1) / 2) * a.mrows * pg.mz + 31) & ~31;
    a.stage_f = CP_TR * pg.pitch;
    a.slotFloats = (long long)NW * S * pg.labelPlane;
    // Ring depth R and block length PB.  A producer of plane p waits for the consumers of the blocks that read plane p-R; their
    // tickets are smaller than its own only if R >= PB + 4 (a block reads PB + 4 planes): that is what rules out deadlock.
    // Default: two blocks with their halo planes, so that the producers of block b+1 seldom wait for the consumers of block b
    // (measured at config 2: NW 3, PB 12, R 28 = a 50 MB ring is the fastest setting).
    const size_t slot_bytes = (size_t)a.slotFloats * sizeof(float);
    a.PB = ca_min(ca_max(env_int("CA_PIPE_PB", 12), 4), 60);
    a.ng = (S + NW - 1) / NW;
    const int nchunks = a.ng * S;
    a.nseq = nchunks * a.nx;
    a.ncolsmax = NW * S * pg.ntC;
    const size_t nctr = 2 + (size_t)a.nseq + (size_t)nchunks * (size_t)(xe - xb) + (size_t)nchunks * a.ncolsmax;
    const size_t ctr_bytes = cp_align256(nctr * sizeof(unsigned));
    // hand-over of the march state between the H blocks of a column (only with T_D inside the producers): 16 float4 per compute
    // thread, double-buffered by block parity and by chunk parity.  Opt-in (CA_PIPE_HANDOFF=1), measured and rejected: it saves
    // the (PB+4)/PB in-plane halo of the consumers but chains the blocks of a column -- a consumer holds its SM slot until its
    // predecessor is done -- 6.89 instead of 6.36 ms per direction at config 2.  Default: every block filters its own halo planes.
    size_t state_bytes = tdp && env_int("CA_PIPE_HANDOFF", 0) != 0 ? cp_align256((size_t)4 * a.ncolsmax * 16 * pg.ncomp * sizeof(float4)) : 0;
    if (ws_bytes < ctr_bytes + state_bytes + 12 * slot_bytes) state_bytes = 0;
    if (ws_bytes < ctr_bytes + 12 * slot_bytes) return CA_ERR_UNSUPPORTED;
    const int r_cap = ca_min(ca_max(env_int("CA_PIPE_R", 2 * a.PB + 4), 12), CP_RMAX);
    a.R = (int)ca_min((int64_t)r_cap, (int64_t)((ws_bytes - ctr_bytes - state_bytes) / slot_bytes));
    if (a.PB > a.R - 4) a.PB = a.R - 4;
    a.nb = (xe - xb + a.PB - 1) / a.PB;
    if (a.nb > CP_MAXB) return CA_ERR_UNSUPPORTED;
    // kW groups of (almost) equal size
    unsigned tick = 0;
    for (int g = 0, k0 = 0; g < a.ng; ++g) {
        const int n = S / a.ng + (g < S % a.ng ? 1 : 0);
        a.nw[g] = n; a.kw0[g] = k0; k0 += n;
        a.tstart[g] = tick;
        const uint64_t per_chunk = (uint64_t)a.nx * pg.ntP + (uint64_t)a.nb * n * S * pg.ntC;
        if ((uint64_t)tick + per_chunk * S >= (1ull << 31)) return CA_ERR_UNSUPPORTED;
        tick += (unsigned)(per_chunk * S);
    }
    a.tstart[a.ng] = tick;
    a.ctr = (unsigned*)ws;
    a.colflag = a.ctr + 2 + (size_t)a.nseq + (size_t)nchunks * (size_t)(xe - xb);
    a.state = state_bytes ? (float*)((char*)ws + ctr_bytes) : nullptr;
    a.ring = (float*)((char*)ws + ctr_bytes + state_bytes);
#ifdef CA_EMU
    a.spin_limit = 1u << 20;
#else
    a.spin_limit = 1u << 23;     // x (64 ns sleep + an L2 round trip) = a few seconds
#endif
    CaTensorMap tmF, tmM;
    int rc = ca_tmap_make(&tmF, fix, d, w, h, C, d, d, CP_TY, (C + 1) / 2);
    if (rc != CA_OK) return rc;
    rc = ca_tmap_make(&tmM, mov, d, w, h, C, d, a.mz, a.mrows, (C + 1) / 2);
    if (rc != CA_OK) return rc;
    void* tmg = ca_timing_begin(CA_T_PIPE, st);
    // counters to zero; ring to zero (its halo rows and pad columns are never written by the kernel)
    if (cudaMemsetAsync(ws, 0, ctr_bytes, st) != cudaSuccess ||
        cudaMemsetAsync(a.ring, 0, (size_t)a.R * slot_bytes, st) != cudaSuccess) {
        ca_set_error("correlate (pipe): memset failed");
        return CA_ERR_LAUNCH;
    }
    g_last_cfg[0] = NW; g_last_cfg[1] = a.PB; g_last_cfg[2] = a.R; g_last_cfg[3] = tdp; g_last_cfg[4] = a.state != nullptr; ++g_last_cfg[5];
    const dim3 grid(tick), block((unsigned)pg.nt);
    CA_SMEM_OPTIN(k, 110 * 1024);
    CA_LAUNCH(k, grid, block, smem, st, tmF, tmM, a);
    ca_timing_end(tmg, st);
    CA_CHECK_LAUNCH("correlate/pipe");
    return CA_OK;
}
}  // namespace

// bytes of workspace the pipe path wants at this shape (0: shape outside its range)
size_t ca_corr_pipe_workspace_bytes(int C, int h, int w, int d, int disp_hw) {
    PipeGeom pg;
    if (!pipe_geom(C, w, d, disp_hw, &pg)) return 0;
    const int S = 2 * disp_hw + 1;
    const size_t nseq = (size_t)S * S * (size_t)h;       // >= nchunks * nx and >= nchunks * nb
    return cp_align256((2 + 2 * nseq + (size_t)S * S * 4 * S * pg.ntC) * sizeof(unsigned)) + (size_t)CP_RMAX * 4 * S * pg.labelPlane * sizeof(float) +
           (size_t)4 * 4 * S * pg.ntC * 16 * pg.ncomp * sizeof(float4) + 512;
}

// fix, mov: unpadded (C,h,w,d) volumes, 16-byte aligned, d % 4 == 0.  Returns CA_ERR_UNSUPPORTED (without an error message
// worth reporting) when the shape is outside the kernel's range or the path is switched off: the caller then takes the
// two-kernel path.
int ca_launch_corr_pipe(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, int xb, int xe, float* ssd,
                        void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!pipe_enabled() || h >= 32768 || w >= 32768) return CA_ERR_UNSUPPORTED;
    if ((((uintptr_t)fix) | ((uintptr_t)mov) | ((uintptr_t)ssd) | ((uintptr_t)ws)) & 15) return CA_ERR_UNSUPPORTED;
    if ((int64_t)(xe - xb) * w * d >= (1ll << 31)) return CA_ERR_UNSUPPORTED;
    int rc;
    switch (disp_hw) {
        case 1: rc = launch_pipe<1>(fix, mov, C, h, w, d, xb, xe, ssd, ws, ws_bytes, st); break;
        case 2: rc = launch_pipe<2>(fix, mov, C, h, w, d, xb, xe, ssd, ws, ws_bytes, st); break;
        case 3: rc = launch_pipe<3>(fix, mov, C, h, w, d, xb, xe, ssd, ws, ws_bytes, st); break;
        case 4: rc = launch_pipe<4>(fix, mov, C, h, w, d, xb, xe, ssd, ws, ws_bytes, st); break;
        case 5: rc = launch_pipe<5>(fix, mov, C, h, w, d, xb, xe, ssd, ws, ws_bytes, st); break;
        case 6: rc = launch_pipe<6>(fix, mov, C, h, w, d, xb, xe, ssd, ws, ws_bytes, st); break;
        case 7: rc = launch_pipe<7>(fix, mov, C, h, w, d, xb, xe, ssd, ws, ws_bytes, st); break;
        case 8: rc = launch_pipe<8>(fix, mov, C, h, w, d, xb, xe, ssd, ws, ws_bytes, st); break;
        default: rc = CA_ERR_UNSUPPORTED; break;
    }
    return rc;
}

// test / debug aid: {NW, PB, R, T_D in the producers, state hand-over, number of launches so far} of the last launch
extern "C" __attribute__((visibility("default"))) void ca_debug_corr_pipe_config(int* out) {
    for (int i = 0; i < 6; ++i) out[i] = g_last_cfg[i];
}

// test / debug aid: non-zero when a hand-off wait of the last launches gave up (see cp_poll)
extern "C" __attribute__((visibility("default"))) int ca_debug_corr_pipe_error(int reset) {
#ifdef CA_EMU
    const unsigned v = g_cp_err;
    if (reset) g_cp_err = 0;
    return (int)v;
#else
    unsigned v = 0, z = 0;
    if (cudaMemcpyFromSymbol(&v, g_cp_err, sizeof(v)) != cudaSuccess) return -1;
    if (reset) cudaMemcpyToSymbol(g_cp_err, &z, sizeof(z));
    return (int)v;
#endif
}
