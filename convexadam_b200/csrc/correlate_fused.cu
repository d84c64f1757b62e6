// correlate() in ONE kernel: raw SSD, both 3x3x3 box filters and the H march fused; nothing but the cost volume touches HBM.
//
// Replaces, for the certified mode (ca_correlate_fast), the pair k_raw_ssd_staged -> chunk buffer -> k_box12_fast of
// correlate.cu, whose raw planes made a round trip through HBM (write 8.6 GB + read 8.6 GB per direction at 192x160x256,
// 3.2x the algorithmic traffic).  Reference: `correlate`, src/convexAdam/convex_adam_utils.py:72-89.
//
// Mathematics.  ssd = box3(box3(raw)) with zero padding and the intermediate truncated to the volume (utils.py:84-85).
// Both boxes and the truncation are separable per axis, so  ssd = T_H T_W T_D raw / 729  with the 1-D operator
// T = B M B (B: zero-padded 3-tap sum, M: mask to [0, n)).  The kernel applies T_D and T_W inside a plane and T_H
// across planes.
//
// Work decomposition.  A CTA owns one (kH,kW) pair -- all S = 2*hw+1 values of kD, i.e. S labels -- and one spatial
// tile of TYO rows x 32 voxels along D, and marches along H (the axis the reference loops over, utils.py:83):
//   stage   : the fixed tile (TYO+4 rows x 40 voxels) of plane X and the moving tile with its displacement
//             neighbourhood (36 + 2*hw voxels) of plane X+kH arrive as two 4-D tensor-map TMA boxes per channel group
//             (cp.async.bulk.tensor, out-of-bounds elements zero-filled = the reference's F.pad, utils.py:76), completion
//             on an mbarrier, NST groups in flight;
//   raw     : thread (row, z-quad) accumulates 4 voxels x S shifts over the channels out of shared memory
//             (packed fp32 pairs: FADD2 + FFMA2), including the 2-voxel halo the filters need;
//   exchange: the raw tile of every label goes through shared memory once (masked to the volume);
//   filter  : thread (label, z-quad, 8 rows) reads 12 rows x 8 voxels, applies T_D and T_W in registers, then T_H with
//             the march state (two previous in-plane results and two previous first-box results: 4 registers per
//             cost-volume entry, 128 per thread) and streams the finished plane X-2 of its 8 rows into `ssd`.
// Per cost-volume entry: 1.41 x (12 FADD2/FFMA2 pairs) for the raw SSD incl. halo, ~7 packed additions for the filters,
// 4 bytes written to HBM; the feature tiles are L2 hits (all CTAs in flight work on neighbouring planes).
//
// Arithmetic: FMA channel sum, separable sums, one multiplication by 1/729 -- at most C + 15 roundings per entry, all
// on sums of non-negative terms, so every entry is within (C + 15) u of the real-number value and, with the reference's
// C + 56 roundings, within (2C + 71) u <= 98 u (C <= 12; 128 u is what the band test of coupled.cu needs) of the
// reference's entry.  The argmin / selections are never taken from this volume directly: see
// ca_coupled_convex_certified.
#include "ca_tma.cuh"

#include <cstdlib>

// 32 output voxels along D per tile, raw tile 2 + 32 + 2.  (Measured alternative: 28 outputs per tile make the raw tile
// exactly 8 quads = four raw warps, one per warp scheduler, for TYO = 12: 21 % fewer cycles per step, but 25 % more tiles
// to cover a 128-voxel row: 7.7 instead of 7.3 ms per direction.)
#define CF_TZO 32   // output voxels along D per tile
#define CF_ZR 36    // raw voxels along D per tile (2 + 32 + 2)
#define CF_NQ 9     // z-quads per raw row
#define CF_OQ 8     // z-quads per output row
#define CF_NST 4    // channel groups in flight
#define CF_CH 6     // channels per staged group

struct CorrFusedArgs {
    float* ssd;
    int C, h, w, d;
    int xb, xe;        // output planes [xb, xe)
    int r0, r1;        // raw planes the outputs need: [xb-2, xe+2) clipped to the volume
    int b0, b1;        // first-box planes: [xb-1, xe+1) clipped to the volume
    int ntz;           // tiles along D
    int NCH;           // channel groups per plane
};

template <int HW, int TYO>
struct CorrFusedGeom {
    static constexpr int S = 2 * HW + 1;
    static constexpr int TYR = TYO + 4;                          // raw rows
    // TMA boxes start on 16-byte boundaries of the global rows: the fixed box starts 2 voxels before the raw range
    // (z0-4), the moving box MSH voxels before the first displaced voxel (z0-2-HW-MSH, MSH = 0 for hw = 2, 6)
    static constexpr int FZ = CF_ZR + 4;                         // fixed-tile row length
    static constexpr int MSH = ((-2 - HW) % 4 + 4) % 4;
    static constexpr int NM = (MSH + 4 + 2 * HW + 3) & ~3;       // moving values loaded per thread (aligned quads)
    static constexpr int MZ = 4 * (CF_NQ - 1) + NM;              // moving-tile row length
    static constexpr int RG = TYO / 4;                           // owner row groups (4 rows each)
    static constexpr int RAW_ITEMS = TYR * CF_NQ;
    static constexpr int OWN_ITEMS = S * CF_OQ * RG;
    static constexpr int RAW_WARPS = (RAW_ITEMS + 31) / 32, OWN_WARPS = (OWN_ITEMS + 31) / 32;
    static constexpr int NT = 32 * (RAW_WARPS + OWN_WARPS + 1);       // + the TMA producer warp
    static constexpr int EX_FLOATS = (S * TYR * CF_ZR + 31) & ~31;     // one exchange buffer
    static constexpr int FIX_CH = TYR * FZ, MOV_CH = TYR * MZ;         // floats per channel
    static constexpr int FIX_F = (CF_CH * FIX_CH + 31) & ~31, STAGE_F = FIX_F + ((CF_CH * MOV_CH + 31) & ~31);
    static constexpr unsigned STAGE_BYTES = (unsigned)(CF_CH * (FIX_CH + MOV_CH)) * 4u;
    static constexpr size_t SMEM = ((size_t)2 * EX_FLOATS + (size_t)CF_NST * STAGE_F) * sizeof(float) + 128;   // + barriers
};

#ifdef CF_PROFILE
// debug build only (tools/build_cf_profile.sh): cycles of raw warp 0 / owner warp 0 per phase, summed over all CTAs
__device__ unsigned long long cf_prof[8];
#define CF_T(var) const long long var = clock64()
#define CF_ACC(slot, t0, t1) prof[slot] += (unsigned long long)((t1) - (t0))
extern "C" __attribute__((visibility("default"))) int ca_debug_cf_prof(unsigned long long* out, int reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyFromSymbol(out, cf_prof, sizeof(z)) != cudaSuccess) return -1;
    if (reset) cudaMemcpyToSymbol(cf_prof, z, sizeof(z));
    return 0;
}
#else
#define CF_T(var)
#define CF_ACC(slot, t0, t1)
#endif

// barriers of one CTA: TMA stage s landed / stage s consumed by every raw warp; exchange buffer e written / read
struct CfBarriers {
    unsigned long long st_full[CF_NST], st_free[CF_NST], ex_full[2], ex_free[2];
};

static __device__ __forceinline__ float4 cf_add3(const float4& a, const float4& b, const float4& c) {
    const float2 lo = ca_add2(ca_add2(make_float2(a.x, a.y), make_float2(b.x, b.y)), make_float2(c.x, c.y));
    const float2 hi = ca_add2(ca_add2(make_float2(a.z, a.w), make_float2(b.z, b.w)), make_float2(c.z, c.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
static __device__ __forceinline__ float4 cf_scale(const float4& a, float s) {
    const float2 lo = ca_mul2(make_float2(a.x, a.y), make_float2(s, s)), hi = ca_mul2(make_float2(a.z, a.w), make_float2(s, s));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// T_D of one exchange row: 8 raw values in (indices 0..7 <-> z = zb-2 .. zb+5), the 4 filtered values of zb .. zb+3 out.
// zm masks the first-box values whose position lies outside [0, d).  DALIGNED (d % 4 == 0): only the two ends can.
template <bool DALIGNED>
static __device__ __forceinline__ float4 cf_td(const float* row, const float (&zm)[6]) {
    const float4 v0 = *reinterpret_cast<const float4*>(row);
    const float4 v1 = *reinterpret_cast<const float4*>(row + 4);
    float a1 = ((v0.x + v0.y) + v0.z) * zm[0];
    float a2 = (v0.y + v0.z) + v0.w;
    float a3 = (v0.z + v0.w) + v1.x;
    float a4 = (v0.w + v1.x) + v1.y;
    float a5 = (v1.x + v1.y) + v1.z;
    float a6 = ((v1.y + v1.z) + v1.w) * zm[5];
    if (!DALIGNED) { a2 *= zm[1]; a3 *= zm[2]; a4 *= zm[3]; a5 *= zm[4]; }
    return make_float4((a1 + a2) + a3, (a2 + a3) + a4, (a3 + a4) + a5, (a4 + a5) + a6);
}

// ---- raw role: thread (row rr, quad rq) of the raw warps ---------------------------------------------------------------
// The thread's 4 voxels x S shifts are accumulated as packed pairs of NEIGHBOURING LABELS of one voxel:
// (m[z+k] - f[z], m[z+k+1] - f[z]) is an aligned register pair of the loaded moving row whenever its first index is even,
// so voxel v pairs the labels (0,1),(2,3).. or (1,2),(3,4).. according to the parity of its first moving index and keeps
// one single label (the last or the first); the fixed value enters as a broadcast operand.  No shuffling of operands.
template <int HW>
struct CfAcc {
    static constexpr int S = 2 * HW + 1;
    float2 p[4][S / 2];     // pairs of voxel v
    float s[4];             // the single label of voxel v
};
// label k of voxel v (PAR = parity of the voxel's first moving index)
template <int HW, int PAR>
static __device__ __forceinline__ float cf_acc_get(const CfAcc<HW>& a, int v, int k) {
    constexpr int S = 2 * HW + 1;
    if (PAR == 0) return k == S - 1 ? a.s[v] : ((k & 1) ? a.p[v][k >> 1].y : a.p[v][k >> 1].x);
    return k == 0 ? a.s[v] : (((k - 1) & 1) ? a.p[v][(k - 1) >> 1].y : a.p[v][(k - 1) >> 1].x);
}

// ---- producer role: one thread keeps CF_NST channel groups in flight ------------------------------------------------------
// group g = (plane X - r0) * NCH + j  ->  channels [6j, 6j+6) of planes X (fixed) and X + kH - hw (moving); channels past C
// are out of bounds of the tensor map and arrive as zeros (they add (0 - 0)^2)
template <int HW, int TYO>
static __device__ __forceinline__ void cf_producer_role(const CaTensorMap* tmF, const CaTensorMap* tmM, const CorrFusedArgs& a,
                                                        float* stages, CfBarriers& bar, int y0, int z0, int kH, int kW) {
    typedef CorrFusedGeom<HW, TYO> G;
    const int NG = (a.r1 - a.r0) * a.NCH;
    for (int g = 0; g < NG; ++g) {
        const int X = a.r0 + g / a.NCH, c0 = (g % a.NCH) * CF_CH, s = g % CF_NST;
        if (g >= CF_NST) ca_mbarx_wait(&bar.st_free[s], (unsigned)(g / CF_NST - 1) & 1u);   // every raw warp has left its previous use
        float* st = stages + (size_t)s * G::STAGE_F;
        ca_mbarx_expect_tx(&bar.st_full[s], G::STAGE_BYTES);
        ca_tma_load_4d(st, tmF, z0 - 4, y0 - 2, X, c0, &bar.st_full[s]);
        ca_tma_load_4d(st + G::FIX_F, tmM, z0 - 2 - HW - G::MSH, y0 - 2 + kW - HW, X + kH - HW, c0, &bar.st_full[s]);
    }
}

template <int HW, int TYO>
static __device__ __forceinline__ void cf_raw_role(const CorrFusedArgs& a, float* ex, float* stages, CfBarriers& bar, int tid, int y0, int z0) {
    typedef CorrFusedGeom<HW, TYO> G;
    constexpr int S = G::S, TYR = G::TYR, NM = G::NM, MZ = G::MZ, FZ = G::FZ, MSH = G::MSH;
    const int lane = tid & 31;
    const bool on = tid < G::RAW_ITEMS;
    const int rr = on ? tid / CF_NQ : 0, rq = on ? tid % CF_NQ : 0;
    bool rzok[4];
    {
        const int y = y0 - 2 + rr;
        const bool rowok = on && y >= 0 && y < a.w;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int z = z0 - 2 + 4 * rq + i;
            rzok[i] = rowok && z >= 0 && z < a.d;
        }
    }
    const int foff = rr * FZ + 4 * rq + 2, moff = rr * MZ + 4 * rq, eoff = rr * CF_ZR + 4 * rq;
    int g = 0;
#ifdef CF_PROFILE
    unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int X = a.r0; X < a.r1; ++X) {
        const int it = X - a.r0;
        CfAcc<HW> acc;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            acc.s[v] = 0.f;
#pragma unroll
            for (int q = 0; q < S / 2; ++q) acc.p[v][q] = make_float2(0.f, 0.f);
        }
        for (int j = 0; j < a.NCH; ++j, ++g) {
            const int s = g % CF_NST;
            const float* fs = stages + (size_t)s * G::STAGE_F + foff;     // 8-byte aligned: two 64-bit loads
            const float* ms = stages + (size_t)s * G::STAGE_F + G::FIX_F + moff;
            CF_T(c0);
            ca_mbarx_wait(&bar.st_full[s], (unsigned)(g / CF_NST) & 1u);
            CF_T(c1);
            CF_ACC(0, c0, c1);
            if (on) {
                // all CF_CH channels of the group, fully unrolled: the loads of channel c+1 are issued before the arithmetic of c
                float2 fa[2], fb[2];
                float4 ma[NM / 4], mb[NM / 4];
                fa[0] = *reinterpret_cast<const float2*>(fs);
                fa[1] = *reinterpret_cast<const float2*>(fs + 2);
#pragma unroll
                for (int q = 0; q < NM / 4; ++q) ma[q] = *reinterpret_cast<const float4*>(ms + 4 * q);
#pragma unroll
                for (int c = 0; c < CF_CH; ++c) {
                    if (c + 1 < CF_CH) {
                        fb[0] = *reinterpret_cast<const float2*>(fs + (c + 1) * G::FIX_CH);
                        fb[1] = *reinterpret_cast<const float2*>(fs + (c + 1) * G::FIX_CH + 2);
#pragma unroll
                        for (int q = 0; q < NM / 4; ++q) mb[q] = *reinterpret_cast<const float4*>(ms + (c + 1) * G::MOV_CH + 4 * q);
                    }
                    const float f[4] = {fa[0].x, fa[0].y, fa[1].x, fa[1].y};
                    float m[NM];
#pragma unroll
                    for (int q = 0; q < NM / 4; ++q) { m[4 * q] = ma[q].x; m[4 * q + 1] = ma[q].y; m[4 * q + 2] = ma[q].z; m[4 * q + 3] = ma[q].w; }
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float2 fv = make_float2(f[v], f[v]);
                        const int par = (MSH + v) & 1;                       // parity of the voxel's first moving index
#pragma unroll
                        for (int q = 0; q < S / 2; ++q) {
                            const int i0 = MSH + v + 2 * q + par;            // even: an aligned pair of the loaded row
                            const float2 df = ca_sub2(make_float2(m[i0], m[i0 + 1]), fv);
                            acc.p[v][q] = ca_fma2(df, df, acc.p[v][q]);
                        }
                        const float ds = m[MSH + v + (par ? 0 : S - 1)] - f[v];
                        acc.s[v] = __fmaf_rn(ds, ds, acc.s[v]);
                    }
                    if (c + 1 < CF_CH) {
                        fa[0] = fb[0]; fa[1] = fb[1];
#pragma unroll
                        for (int q = 0; q < NM / 4; ++q) ma[q] = mb[q];
                    }
                }
            }
            // this warp is done with stage s: the producer warp refills it once every raw warp has said so
            CF_T(c2);
            CF_ACC(1, c1, c2);
            ca_mbarx_arrive(&bar.st_free[s]);               // every thread arrives for its own reads (release)
            CF_T(c3);
            CF_ACC(2, c2, c3);
        }
        // hand the raw tile of plane X to the owners through exchange buffer it & 1
        const int e = it & 1;
        CF_T(c4);
        if (it >= 2) ca_mbarx_wait(&bar.ex_free[e], (unsigned)((it >> 1) - 1) & 1u);
        CF_T(c5);
        CF_ACC(3, c4, c5);
        if (on) {
            float* exw = ex + e * G::EX_FLOATS + eoff;
#pragma unroll
            for (int k = 0; k < S; ++k) {
                float4 o;
                o.x = rzok[0] ? cf_acc_get<HW, (MSH + 0) & 1>(acc, 0, k) : 0.f;
                o.y = rzok[1] ? cf_acc_get<HW, (MSH + 1) & 1>(acc, 1, k) : 0.f;
                o.z = rzok[2] ? cf_acc_get<HW, (MSH + 2) & 1>(acc, 2, k) : 0.f;
                o.w = rzok[3] ? cf_acc_get<HW, (MSH + 3) & 1>(acc, 3, k) : 0.f;
                *reinterpret_cast<float4*>(exw + k * (TYR * CF_ZR)) = o;
            }
        }
        ca_mbarx_arrive(&bar.ex_full[e]);                   // every thread arrives for its own writes (release)
        CF_T(c6);
        CF_ACC(4, c5, c6);
    }
#ifdef CF_PROFILE
    if (tid == 0)
        for (int i = 0; i < 5; ++i) atomicAdd(&cf_prof[i], prof[i]);
#endif
}

// ---- owner role: thread (label k, quad zq, 4 rows) of the owner warps ---------------------------------------------------
// One H step for the thread's 4 rows.  P = parity of the step: state slot P holds the older of the two previous values.
template <int P, bool DALIGNED>
static __device__ __forceinline__ void cf_owner_step(const float* exr, bool have_raw, bool p1ok, bool store, float4 (&hq)[4][2],
                                                     float4 (&ha)[4][2], const float (&zm)[6], const float (&ym)[6], float* outp, int d,
                                                     int nrows, int nz) {
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float c729 = 1.0f / 729.0f;
    // exchange rows i = 0..7 <-> y = oy - 2 + i; first box along W on rows 1..6, second on rows 2..5 = the 4 outputs
    float4 dm, d0, bm, b0;
    if (have_raw) { dm = cf_td<DALIGNED>(exr, zm); d0 = cf_td<DALIGNED>(exr + CF_ZR, zm); }
    else { dm = d0 = zero4; }
    bm = b0 = zero4;
#pragma unroll
    for (int i = 1; i <= 6; ++i) {
        const float4 d1 = have_raw ? cf_td<DALIGNED>(exr + (i + 1) * CF_ZR, zm) : zero4;
        const float4 b1 = cf_scale(cf_add3(dm, d0, d1), ym[i - 1]);
        dm = d0; d0 = d1;
        if (i >= 3) {
            const int lr = i - 3;                                           // output row: e = b_{i-2} + b_{i-1} + b_i
            const float4 in = cf_add3(bm, b0, b1);
            float4 A = cf_add3(hq[lr][P], hq[lr][1 - P], in);                // first box along H: plane X-1 ...
            if (!p1ok) A = zero4;                                           // ... outside the volume: truncated
            hq[lr][P] = in;
            const float4 o = cf_scale(cf_add3(ha[lr][P], ha[lr][1 - P], A), c729);   // second box along H: plane X-2
            ha[lr][P] = A;
            if (store && lr < nrows) {
                float* p = outp + (int64_t)lr * d;
                if (DALIGNED) {
                    if (nz > 0) ca_st_stream4(reinterpret_cast<float4*>(p), o);
                } else {
                    if (nz > 0) p[0] = o.x;
                    if (nz > 1) p[1] = o.y;
                    if (nz > 2) p[2] = o.z;
                    if (nz > 3) p[3] = o.w;
                }
            }
        }
        bm = b0; b0 = b1;
    }
}

template <int HW, int TYO, bool DALIGNED>
static __device__ __forceinline__ void cf_owner_role(const CorrFusedArgs& a, const float* ex, CfBarriers& bar, int otid, int y0, int z0,
                                                     int pair) {
    typedef CorrFusedGeom<HW, TYO> G;
    constexpr int S = G::S, TYR = G::TYR;
    const int lane = otid & 31;
    const bool on = otid < G::OWN_ITEMS;
    const int zq = otid % CF_OQ, rg = (otid / CF_OQ) % G::RG, k = on ? (otid / CF_OQ) / G::RG : 0;
    const int oy = y0 + 4 * rg, oz = z0 + 4 * zq;
    float zm[6], ym[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const int z = oz - 1 + j;                                           // first-box position along D
        zm[j] = (z >= 0 && z < a.d) ? 1.f : 0.f;
        const int y = oy - 1 + j;                                           // first-box row
        ym[j] = (y >= 0 && y < a.w) ? 1.f : 0.f;
    }
    float4 hq[4][2], ha[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int p = 0; p < 2; ++p) hq[r][p] = ha[r][p] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nx = a.xe - a.xb;
    const int nrows = on ? a.w - oy : 0, nz = a.d - oz;                      // rows / voxels of the thread inside the volume
    const int64_t label = (int64_t)k * S * S + pair;
    float* outp = a.ssd + ((label * nx) * a.w + oy) * (int64_t)a.d + oz;     // plane xb of the thread's first row
    const int64_t outPitch = (int64_t)a.w * a.d;
    const int eoff = (k * TYR + 4 * rg) * CF_ZR + 4 * zq;
#ifdef CF_PROFILE
    unsigned long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (int X = a.r0; X <= a.xe + 1; ++X) {
        const int it = X - a.r0, e = it & 1;
        const bool have_raw = X < a.r1;
        const bool p1ok = X - 1 >= a.b0 && X - 1 < a.b1;
        const bool store = X - 2 >= a.xb;
        CF_T(c0);
        if (have_raw) ca_mbarx_wait(&bar.ex_full[e], (unsigned)(it >> 1) & 1u);
        CF_T(c1);
        CF_ACC(5, c0, c1);
        const float* exr = ex + e * G::EX_FLOATS + eoff;
        if (it & 1) cf_owner_step<1, DALIGNED>(exr, have_raw, p1ok, store, hq, ha, zm, ym, outp, a.d, nrows, nz);
        else cf_owner_step<0, DALIGNED>(exr, have_raw, p1ok, store, hq, ha, zm, ym, outp, a.d, nrows, nz);
        if (have_raw) ca_mbarx_arrive(&bar.ex_free[e]);     // every thread arrives for its own reads
        if (store) outp += outPitch;
    }
}

template <int HW, int TYO, bool DALIGNED>
__global__ void __launch_bounds__((CorrFusedGeom<HW, TYO>::NT), 1)
    k_corr_fused(CA_TMAP_PARAM tmF, CA_TMAP_PARAM tmM, CorrFusedArgs a) {
    typedef CorrFusedGeom<HW, TYO> G;
    // tensor-map TMA writes need 128-byte aligned destinations: the dynamic window is declared with that alignment (and
    // nothing static precedes it -- the barriers live at its end), so every pointer below keeps its shared address space
    CA_DYN_SMEM_ALIGNED(float, sm, 128);
    float* ex = sm;                                                       // [2][S][TYR][CF_ZR]
    float* stages = sm + 2 * G::EX_FLOATS;                                // [NST][CH x (fixed tile | moving tile)]
    CfBarriers& bar = *reinterpret_cast<CfBarriers*>(stages + CF_NST * G::STAGE_F);
    const int tid = (int)threadIdx.x;
    const int tz = (int)blockIdx.x % a.ntz, ty = (int)blockIdx.x / a.ntz;
    const int y0 = ty * TYO, z0 = tz * CF_TZO;
    const int pair = (int)blockIdx.y;
    if (tid == 0) {
        for (int s = 0; s < CF_NST; ++s) { ca_mbarx_init(&bar.st_full[s], 1); ca_mbarx_init(&bar.st_free[s], 32 * G::RAW_WARPS); }
        for (int e = 0; e < 2; ++e) { ca_mbarx_init(&bar.ex_full[e], 32 * G::RAW_WARPS); ca_mbarx_init(&bar.ex_free[e], 32 * G::OWN_WARPS); }
    }
    __syncthreads();
    if (tid < 32 * G::RAW_WARPS) cf_raw_role<HW, TYO>(a, ex, stages, bar, tid, y0, z0);
    else if (tid < 32 * (G::RAW_WARPS + G::OWN_WARPS)) cf_owner_role<HW, TYO, DALIGNED>(a, ex, bar, tid - 32 * G::RAW_WARPS, y0, z0, pair);
    else if (tid == 32 * (G::RAW_WARPS + G::OWN_WARPS)) cf_producer_role<HW, TYO>(&tmF, &tmM, a, stages, bar, y0, z0, pair % G::S, pair / G::S);
}

// ---- host side ---------------------------------------------------------------------------------------------------------
namespace {
// CA_CORR_FUSED=1 switches the certified mode to this kernel.  Off by default: at 192x160x256 it measures 7.8 ms per
// direction against the two-kernel path's time (DESIGN.md 4.1 has the numbers and the reason: the kernel is bound by the
// fp32 FMA lanes, and pays 1.5x the raw arithmetic for its halo).  Read on every call so that tests can toggle it.
bool fused_enabled() {
    const char* e = getenv("CA_CORR_FUSED");
    return e && atoi(e) == 1;
}

template <int HW, int TYO>
int launch_fused(const float* fix, const float* mov, int C, int h, int w, int d, int pitch, int xb, int xe, float* ssd,
                 cudaStream_t st) {
    typedef CorrFusedGeom<HW, TYO> G;
    CorrFusedArgs a;
    a.ssd = ssd; a.C = C; a.h = h; a.w = w; a.d = d;
    a.xb = xb; a.xe = xe;
    a.r0 = ca_max(0, xb - 2); a.r1 = ca_min(h, xe + 2);
    a.b0 = ca_max(0, xb - 1); a.b1 = ca_min(h, xe + 1);
    a.ntz = (d + CF_TZO - 1) / CF_TZO;
    a.NCH = (C + CF_CH - 1) / CF_CH;
    const bool aligned = (d & 3) == 0 && (((uintptr_t)ssd) & 15) == 0;
    const size_t smem = G::SMEM;
    if (smem > 227 * 1024) {
        ca_set_error("correlate (fused): tile needs %zu bytes of shared memory", smem);
        return CA_ERR_UNSUPPORTED;
    }
    CaTensorMap tmF, tmM;
    int rc = ca_tmap_make(&tmF, fix, d, w, h, C, pitch, G::FZ, G::TYR, CF_CH);
    if (rc != CA_OK) return rc;
    rc = ca_tmap_make(&tmM, mov, d, w, h, C, pitch, G::MZ, G::TYR, CF_CH);
    if (rc != CA_OK) return rc;
    const int nty = (w + TYO - 1) / TYO;
    const dim3 grid((unsigned)(a.ntz * nty), (unsigned)(G::S * G::S)), block((unsigned)G::NT);
    if (aligned) { auto k = k_corr_fused<HW, TYO, true>; CA_SMEM_OPTIN(k, smem); CA_LAUNCH(k, grid, block, smem, st, tmF, tmM, a); }
    else { auto k = k_corr_fused<HW, TYO, false>; CA_SMEM_OPTIN(k, smem); CA_LAUNCH(k, grid, block, smem, st, tmF, tmM, a); }
    CA_CHECK_LAUNCH("correlate/fused");
    return CA_OK;
}
}  // namespace

// fix, mov: (C,h,w,pitch) with rows of `pitch` floats (pitch % 4 == 0, 16-byte aligned bases), d valid voxels per row.
// Returns CA_ERR_UNSUPPORTED (without an error message worth reporting) when the shape is outside the kernel's range:
// the caller then takes the two-kernel path.
int ca_launch_corr_fused(const float* fix, const float* mov, int C, int h, int w, int d, int pitch, int disp_hw, int xb, int xe,
                         float* ssd, cudaStream_t st) {
    if (!fused_enabled() || disp_hw > 6 || C > 24 || h >= 32768 || w >= 32768) return CA_ERR_UNSUPPORTED;
    if ((int64_t)(xe - xb) * w * d >= (1ll << 31)) return CA_ERR_UNSUPPORTED;
    void* tm = ca_timing_begin(CA_T_FUSED, st);
    int rc;
    switch (disp_hw) {
        case 1: rc = launch_fused<1, 12>(fix, mov, C, h, w, d, pitch, xb, xe, ssd, st); break;
        case 2: rc = launch_fused<2, 12>(fix, mov, C, h, w, d, pitch, xb, xe, ssd, st); break;
        case 3: rc = launch_fused<3, 12>(fix, mov, C, h, w, d, pitch, xb, xe, ssd, st); break;
        case 4: rc = launch_fused<4, 12>(fix, mov, C, h, w, d, pitch, xb, xe, ssd, st); break;
        case 5: rc = launch_fused<5, 12>(fix, mov, C, h, w, d, pitch, xb, xe, ssd, st); break;
        default: rc = launch_fused<6, 12>(fix, mov, C, h, w, d, pitch, xb, xe, ssd, st); break;
    }
    ca_timing_end(tm, st);
    return rc;
}
