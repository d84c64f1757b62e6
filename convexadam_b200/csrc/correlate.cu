// Dense SSD correlation volume over the (2*hw+1)^3 displacement search, with the two 3x3x3 box filters.
//
// Replaces `correlate` (reference src/convexAdam/convex_adam_utils.py:72-89).  The reference materialises an
// im2col of the zero-padded moving features (utils.py:76-77, 9 GB at 192x160x256) and loops over the H-shift;
// here nothing of that exists: labels are produced in chunks of G (kH,kW) pairs x all kD by two kernels,
//     raw   (k_raw_ssd) : per thread 4 consecutive voxels along D x all S values of kD, channel loop outside
//                         -> every feature load is a 128-bit vector load reused S times from registers
//     box12 (k_box12)   : BOTH zero-padded 27-tap box filters fused.  A CTA owns one label plane x a tile of rows
//                         and marches along H; every thread keeps, for 2 rows x 4 voxels, the partial sums of the
//                         three outputs (x-1, x, x+1) a plane contributes to, so each input value is read once from
//                         memory per plane and used for 27 additions out of registers (the +-1 neighbours along D
//                         come from warp shuffles).  The first box's plane goes through a double-buffered
//                         shared-memory tile to the second box, whose finished plane streams into the (L,h,w,d)
//                         volume.  Sums are the reference's: ascending (dH,dW,dD), one accumulator, then /27.
// The kernels are fp32-issue-bound by construction (94 dependent-order flops per cost-volume entry, no FMA
// contraction allowed), not HBM-bound: see DESIGN.md for the arithmetic floor.
//
// EXACT mode: this is a bit-exact restatement of the torch-CPU arithmetic -- (f-m) rounded, squared (x*x)
// rounded, channels accumulated sequentially from c=0, each box3 a sequential fp32 sum of the in-volume taps
// followed by a true division by 27 (avg_pool3d, count_include_pad) -- so that the argmin over labels is the
// reference's argmin.  Every arithmetic op is an explicit _rn intrinsic: nvcc must not contract them to FMAs.
#include "box_march.cuh"
#include "ca_tma.cuh"

#include <cstdlib>

int ca_launch_cost_argmin(const float* ssd, int64_t n, int L, const float* mesh, const float* soft, float coeff,
                          unsigned long long* keys, cudaStream_t st);
int ca_launch_keys_to_i64(const unsigned long long* keys, int64_t n, int64_t* out, cudaStream_t st);
// correlate_fused.cu: the whole stage in one kernel (certified mode); CA_ERR_UNSUPPORTED = take the two-kernel path
int ca_launch_corr_fused(const float* fix, const float* mov, int C, int h, int w, int d, int pitch, int disp_hw, int xb, int xe,
                         float* ssd, cudaStream_t st);
// correlate_pipe.cu: producer / consumer CTAs of one launch exchanging the raw planes through a ring in L2 (certified mode);
// `ws` = scratch for the ring and its counters; CA_ERR_UNSUPPORTED = take the two-kernel path
int ca_launch_corr_pipe(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, int xb, int xe, float* ssd,
                        void* ws, size_t ws_bytes, cudaStream_t st);

// out (C, h+2p, w+2p, pitch) <- in (C,h,w,d), zero elsewhere (F.pad at utils.py:76 plus 16-byte row alignment)
__global__ void __launch_bounds__(256) k_pad_features(const float* __restrict__ in, int C, int h, int w, int d, int p,
                                                      int pitch, float* __restrict__ out) {
    const int hp = h + 2 * p, wp = w + 2 * p;
    const int64_t total = (int64_t)C * hp * wp * pitch;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(o % pitch) - p;
        const int y = (int)((o / pitch) % wp) - p;
        const int x = (int)((o / ((int64_t)pitch * wp)) % hp) - p;
        const int c = (int)(o / ((int64_t)pitch * wp * hp));
        float v = 0.f;
        if (x >= 0 && x < h && y >= 0 && y < w && z >= 0 && z < d) v = __ldg(in + (((int64_t)c * h + x) * w + y) * d + z);
        out[o] = v;
    }
}

// zero halo of the chunk buffer (rows -2,-1,w,w+1 of every plane slice and the 4+4 pad columns of the others): the
// raw kernels only ever write interiors, the box kernels read rows/columns around a tile without bounds checks
__global__ void __launch_bounds__(128) k_zero_raw_halo(float* __restrict__ raw, int64_t nslices, int w, int rowPitch) {
    const int rq = rowPitch / 4;
    const int per_slice = 4 * rq + 2 * w;                      // float4 items: 4 whole rows + 2 pad quads per interior row
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        float* base = raw + sl * (int64_t)(w + 4) * rowPitch;
        for (int k = (int)threadIdx.x; k < per_slice; k += (int)blockDim.x) {
            if (k < 4 * rq) {
                const int r = k / rq, q = k - r * rq;
                const int row = r < 2 ? r : w + r;             // rows 0, 1, w+2, w+3
                *reinterpret_cast<float4*>(base + row * rowPitch + 4 * q) = z;
            } else {
                const int j = k - 4 * rq, row = 2 + (j >> 1);
                *reinterpret_cast<float4*>(base + row * rowPitch + ((j & 1) ? rowPitch - 4 : 0)) = z;
            }
        }
    }
}

// ---- tile geometry of the raw-SSD kernel ---------------------------------------------------------------------------
// A block is (bx, by) threads: bx consecutive z-quads (4 voxels each) x by consecutive y rows; grid = (tiles, h, planes)
// so that no thread ever needs a 64-bit div/mod (the kernels are issue-bound, index maths must be free).
struct TileGeom {
    int zq, nzb, bx, by;
};
static TileGeom make_tile(int dF, int nthreads) {
    TileGeom t;
    t.zq = dF >> 2;
    t.bx = 1;
    while (t.bx < t.zq && t.bx < 32) t.bx <<= 1;
    t.by = nthreads / t.bx;
    t.nzb = (t.zq + t.bx - 1) / t.bx;
    return t;
}

// Chunk buffer of raw SSD planes: (G*S planes) x nx x (w+4) x (dF+8) floats, voxel (x,y,z) of a plane at
// [x-r0][y+2][z+4], where [r0, r0+nx) are the H planes a slab of outputs needs (the whole volume: r0 = 0, nx = h).  Rows -2,-1,w,w+1 and the 4+4 pad columns are zero (memset once per call, kernels write interiors
// only), columns d..dF-1 are written as zero: the box kernel reads rows/columns around a tile without bounds checks.
struct RawLayout {
    int rowPitch;       // dF + 8
    int64_t xPitch;     // (w + 4) * rowPitch
    int64_t plane;      // nx * xPitch
};
static RawLayout make_raw_layout(int nx, int w, int dF) {
    RawLayout r;
    r.rowPitch = dF + 8;
    r.xPitch = (int64_t)(w + 4) * r.rowPitch;
    r.plane = (int64_t)nx * r.xPitch;
    return r;
}

// raw[(g*S + kD), x, y, z] = sum_c (f_c(x,y,z) - mpad_c(x+kH, y+kW, z+kD))^2 for the pair (kH,kW) = chunk_start+g
// FAST: the channel sum is an FMA chain (14 roundings per value instead of the reference's 36 separate ops); see the
// "certified" mode below for how its results are used.
template <int HW, int MINB, bool FAST>
__global__ void __launch_bounds__(128, MINB) k_raw_ssd(const float* __restrict__ fixp, const float* __restrict__ movp,
                                                       int C, int h, int w, int d, int dF, int dP, int chunk_start,
                                                       int zq, int nzb, RawLayout rl, int r0, float* __restrict__ raw) {
    constexpr int S = 2 * HW + 1;
    constexpr int NM = (4 + 2 * HW + 3) & ~3;   // moving values touched by 4 voxels x S shifts, rounded to float4
    // grid = (pairs of the chunk, tiles of a plane, h): the blocks in flight share one slab of the feature volumes
    // (same x, all (kH,kW) shifts), so features come from L2, not HBM
    const int bz = (int)blockIdx.y % nzb, byi = (int)blockIdx.y / nzb;
    const int q = bz * (int)blockDim.x + (int)threadIdx.x;
    const int y = byi * (int)blockDim.y + (int)threadIdx.y;
    const int x = r0 + (int)blockIdx.z;
    if (q >= zq || y >= w) return;
    const int z0 = q * 4;
    const int pi = chunk_start + (int)blockIdx.x;
    const int kH = pi % S, kW = pi / S;
    const int hp = h + 2 * HW, wp = w + 2 * HW;

    float acc[S][4];
#pragma unroll
    for (int k = 0; k < S; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;

    const float* fp = fixp + ((int64_t)x * w + y) * dF + z0;
    const float* mp = movp + ((int64_t)(x + kH) * wp + (y + kW)) * dP + z0;
    const int64_t fstride = (int64_t)h * w * dF, mstride = (int64_t)hp * wp * dP;
    // software pipeline: the loads of channel c+1 are in flight while channel c is reduced (the kernel is bound by
    // the latency of these L2 hits, not by issue slots)
    float4 f4 = __ldg(reinterpret_cast<const float4*>(fp));
    float4 m4[NM / 4];
#pragma unroll
    for (int j = 0; j < NM / 4; ++j) m4[j] = __ldg(reinterpret_cast<const float4*>(mp) + j);
    for (int c = 0; c < C; ++c) {
        const float f[4] = {f4.x, f4.y, f4.z, f4.w};
        float m[NM];
#pragma unroll
        for (int j = 0; j < NM / 4; ++j) { m[4 * j + 0] = m4[j].x; m[4 * j + 1] = m4[j].y; m[4 * j + 2] = m4[j].z; m[4 * j + 3] = m4[j].w; }
        if (c + 1 < C) {
            fp += fstride;
            mp += mstride;
            f4 = __ldg(reinterpret_cast<const float4*>(fp));
#pragma unroll
            for (int j = 0; j < NM / 4; ++j) m4[j] = __ldg(reinterpret_cast<const float4*>(mp) + j);
        }
#pragma unroll
        for (int k = 0; k < S; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float df = __fsub_rn(f[i], m[i + k]);
                acc[k][i] = FAST ? __fmaf_rn(df, df, acc[k][i]) : __fadd_rn(acc[k][i], __fmul_rn(df, df));
            }
    }
    float* rp = raw + (int64_t)blockIdx.x * S * rl.plane + (int64_t)blockIdx.z * rl.xPitch + (int64_t)(y + 2) * rl.rowPitch + 4 + z0;
#pragma unroll
    for (int k = 0; k < S; ++k) {
        float4 o;
        o.x = (z0 + 0 < d) ? acc[k][0] : 0.f;
        o.y = (z0 + 1 < d) ? acc[k][1] : 0.f;
        o.z = (z0 + 2 < d) ? acc[k][2] : 0.f;
        o.w = (z0 + 3 < d) ? acc[k][3] : 0.f;
        ca_st_stream4(reinterpret_cast<float4*>(rp + k * rl.plane), o);
    }
}

// ---- raw SSD, staged: the kernel the full-size path runs -------------------------------------------------------------
// k_raw_ssd above reads every feature value straight from L2 and is bound by the latency of those loads (ncu: 63 % of
// the stall samples sit on the first use of a loaded value, the L1 miss queues are full).  Here a CTA owns TY whole rows
// of one plane x and one H-shift kH, and produces ALL S values of kD for NW consecutive values of kW from shared
// memory: per channel the fixed rows (TY x dF, contiguous in memory) and the moving rows with their displacement
// neighbourhood ((TY+NW-1) x dP, contiguous too: the features are stored zero-padded) arrive as ONE bulk asynchronous
// copy each (cp.async.bulk -> UBLKCP, completion on an mbarrier): 2*C copies per CTA, the SM's load/store units and L1
// are out of the picture, and each feature value is fetched from L2 once per CTA instead of once per (kW, thread).
// The arithmetic per thread is that of k_raw_ssd.
#define CA_RAW_TY 4
struct RawStagedArgs {
    const float* fixp;
    const float* movp;
    float* raw;
    RawLayout rl;
    int C, h, w, d, dF, dP, r0, kW0, NW;
};

template <int HW, bool FAST>
__global__ void __launch_bounds__(256, 2) k_raw_ssd_staged(RawStagedArgs a) {
    constexpr int S = 2 * HW + 1;
    constexpr int NM = (4 + 2 * HW + 3) & ~3;
    CA_DYN_SMEM(float, sm);
    __shared__ unsigned long long bar;
    const int nthr = (int)(blockDim.x * blockDim.y);
    const int tid = (int)threadIdx.y * (int)blockDim.x + (int)threadIdx.x;
    const int kH = (int)blockIdx.x;
    const int x = a.r0 + (int)blockIdx.z;
    const int y0t = (int)blockIdx.y * CA_RAW_TY;
    const int hp = a.h + 2 * HW, wp = a.w + 2 * HW;
    const int nfr = ca_min(CA_RAW_TY, a.w - y0t);                    // fixed rows of this tile
    const int nmr = nfr + a.NW - 1;                                  // moving rows: y + kW
    const int mrows = CA_RAW_TY + a.NW - 1;
    const int fch = CA_RAW_TY * a.dF, mch = mrows * a.dP;            // floats per channel in shared memory
    float* sfix = sm;                                                // [C][TY][dF]
    float* smov = sm + (size_t)a.C * fch;                            // [C][TY+NW-1][dP]
    if (tid == 0) ca_mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) ca_mbar_expect_tx(&bar, (unsigned)(a.C * (nfr * a.dF + nmr * a.dP) * 4));
    for (int i = tid; i < 2 * a.C; i += nthr) {
        const int c = i >> 1;
        if (i & 1) ca_bulk_g2s(smov + (size_t)c * mch, a.movp + (((int64_t)c * hp + x + kH) * wp + y0t + a.kW0) * a.dP, (unsigned)(nmr * a.dP) * 4, &bar);
        else ca_bulk_g2s(sfix + (size_t)c * fch, a.fixp + (((int64_t)c * a.h + x) * a.w + y0t) * a.dF, (unsigned)(nfr * a.dF) * 4, &bar);
    }
    ca_mbar_wait(&bar, 0);
    const int ty = (int)threadIdx.y, z0 = 4 * (int)threadIdx.x;
    const int y = y0t + ty;
    const bool active = y < a.w;
    // Interior rows only.  The 4+4 zero pad columns of a row (read by the box kernels without bounds checks) are written
    // here by the row's edge threads; the halo ROWS (-2,-1,w,w+1) are never written: the fast box kernel does not copy
    // them (its stages stay zero there) and the exact path zeroes them once per call (k_zero_raw_halo).  Round 1 wrote
    // the halo rows from here too: those predicated stores were half of the kernel's executed instructions.
    const bool lpad = threadIdx.x == 0, rpad = threadIdx.x == blockDim.x - 1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* fs = sfix + ty * a.dF + z0;
    float* rp0 = a.raw + (int64_t)blockIdx.z * a.rl.xPitch + 4 + z0 + (int64_t)(y + 2) * a.rl.rowPitch;
    if (!active) return;
    for (int kWl = 0; kWl < a.NW; ++kWl) {
        float acc[S][4];
        const float* ms = smov + (ty + kWl) * a.dP + z0;
        if (FAST) {
            // packed pairs of NEIGHBOURING LABELS of one voxel: (m[z+k] - f[z], m[z+k+1] - f[z]) is an aligned register pair of
            // the loaded moving row whenever z+k is even, so voxel v pairs the labels (0,1),(2,3).. (v even) or (1,2),(3,4)..
            // (v odd) and keeps one single label; the fixed value is a broadcast operand.  No operand shuffling.
            float2 pr[4][S / 2];
            float sg[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                sg[v] = 0.f;
#pragma unroll
                for (int q = 0; q < S / 2; ++q) pr[v][q] = make_float2(0.f, 0.f);
            }
#pragma unroll 2
            for (int c = 0; c < a.C; ++c) {
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
                    const int par = v & 1;
#pragma unroll
                    for (int q = 0; q < S / 2; ++q) {
                        const int i0 = v + 2 * q + par;
                        const float2 df = ca_sub2(make_float2(m[i0], m[i0 + 1]), fv);
                        pr[v][q] = ca_fma2(df, df, pr[v][q]);
                    }
                    const float ds = m[v + (par ? 0 : S - 1)] - f[v];
                    sg[v] = __fmaf_rn(ds, ds, sg[v]);
                }
            }
#pragma unroll
            for (int k = 0; k < S; ++k)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if ((v & 1) == 0) acc[k][v] = k == S - 1 ? sg[v] : ((k & 1) ? pr[v][k >> 1].y : pr[v][k >> 1].x);
                    else acc[k][v] = k == 0 ? sg[v] : (((k - 1) & 1) ? pr[v][(k - 1) >> 1].y : pr[v][(k - 1) >> 1].x);
                }
        } else {
            // exact mode stays scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (even through an fma with a
            // -0 addend), which it never does for the scalar _rn intrinsics
#pragma unroll
            for (int k = 0; k < S; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;
            for (int c = 0; c < a.C; ++c) {
                const float4 f4 = *reinterpret_cast<const float4*>(fs + c * fch);
                const float f[4] = {f4.x, f4.y, f4.z, f4.w};
                float m[NM];
#pragma unroll
                for (int j = 0; j < NM / 4; ++j) {
                    const float4 m4 = *reinterpret_cast<const float4*>(ms + c * mch + 4 * j);
                    m[4 * j + 0] = m4.x; m[4 * j + 1] = m4.y; m[4 * j + 2] = m4.z; m[4 * j + 3] = m4.w;
                }
#pragma unroll
                for (int k = 0; k < S; ++k)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float df = __fsub_rn(f[i], m[i + k]);
                        acc[k][i] = __fadd_rn(acc[k][i], __fmul_rn(df, df));
                    }
            }
        }
        float* rp = rp0 + (int64_t)(kWl * S + kH) * S * a.rl.plane;
#pragma unroll
        for (int k = 0; k < S; ++k) {
            float4 o;
            o.x = (z0 + 0 < a.d) ? acc[k][0] : 0.f;
            o.y = (z0 + 1 < a.d) ? acc[k][1] : 0.f;
            o.z = (z0 + 2 < a.d) ? acc[k][2] : 0.f;
            o.w = (z0 + 3 < a.d) ? acc[k][3] : 0.f;
            ca_st_stream4(reinterpret_cast<float4*>(rp + k * a.rl.plane), o);
            if (lpad) ca_st_stream4(reinterpret_cast<float4*>(rp + k * a.rl.plane - 4), zero4);
            if (rpad) ca_st_stream4(reinterpret_cast<float4*>(rp + k * a.rl.plane + 4), zero4);
        }
    }
}

// ---- raw SSD staged by tensor-map TMA: the certified mode's default -----------------------------------------------------
// Same tile, thread mapping and arithmetic as k_raw_ssd_staged<HW, true>, but the two tiles of a CTA arrive as ONE 4-D
// tensor-map box each (cp.async.bulk.tensor.4d -> UTMALDG.4D) straight from the UNPADDED (C,h,w,d) feature volumes: the
// fixed box (d, TY, 1, C) at (0, y0, x, 0) and the moving box (d + 2*hw rounded up, TY+NW-1, 1, C) at
// (-hw-MSH, y0+kW0-hw, x+kH-hw, 0).  Elements outside the tensor are filled with zeros by the TMA unit -- the reference's
// F.pad(mind_mov, hw) (utils.py:76) and the ragged last row tile at once -- so the zero-padded copy of the moving features
// (k_pad_features: 68 MB written and re-read per direction at config 2) is gone.  A box must start on a 16-byte boundary of
// the row, hence the MSH = (-hw) mod 4 extra leading voxels, skipped by a compile-time register offset.
struct RawTmapArgs {
    float* raw;
    RawLayout rl;
    int C, w, d, r0, kW0, NW;
    int fix_f;          // floats of the fixed tile, rounded up to a 128-byte multiple (offset of the moving tile)
    int mz;             // moving-tile row length
};

template <int HW>
__global__ void __launch_bounds__(256, 2) k_raw_ssd_tmap(CA_TMAP_PARAM tmF, CA_TMAP_PARAM tmM, RawTmapArgs a) {
    constexpr int S = 2 * HW + 1;
    constexpr int MSH = (4 - HW % 4) % 4;
    constexpr int NM = (MSH + 4 + 2 * HW + 3) & ~3;
    CA_DYN_SMEM_ALIGNED(float, sm, 128);
    const int tid = (int)threadIdx.y * (int)blockDim.x + (int)threadIdx.x;
    const int kH = (int)blockIdx.x;
    const int x = a.r0 + (int)blockIdx.z;
    const int y0t = (int)blockIdx.y * CA_RAW_TY;
    const int mrows = CA_RAW_TY + a.NW - 1;
    const int fch = CA_RAW_TY * a.d, mch = mrows * a.mz;             // floats per channel in shared memory
    float* sfix = sm;                                                // [C][TY][d]
    float* smov = sm + a.fix_f;                                      // [C][TY+NW-1][mz]
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smov + (size_t)a.C * mch);
    if (tid == 0) ca_mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        ca_mbar_expect_tx(bar, (unsigned)(a.C * (fch + mch)) * 4u);
        ca_tma_load_4d(sfix, &tmF, 0, y0t, x, 0, bar);
        ca_tma_load_4d(smov, &tmM, -HW - MSH, y0t + a.kW0 - HW, x + kH - HW, 0, bar);
    }
    ca_mbar_wait(bar, 0);
    const int ty = (int)threadIdx.y, z0 = 4 * (int)threadIdx.x;
    const int y = y0t + ty;
    if (y >= a.w) return;
    const bool lpad = threadIdx.x == 0, rpad = threadIdx.x == blockDim.x - 1;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* fs = sfix + ty * a.d + z0;
    float* rp0 = a.raw + (int64_t)blockIdx.z * a.rl.xPitch + 4 + z0 + (int64_t)(y + 2) * a.rl.rowPitch;
    for (int kWl = 0; kWl < a.NW; ++kWl) {
        const float* ms = smov + (ty + kWl) * a.mz + z0;
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
#pragma unroll 2
        for (int c = 0; c < a.C; ++c) {
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
        float* rp = rp0 + (int64_t)(kWl * S + kH) * S * a.rl.plane;
#pragma unroll
        for (int k = 0; k < S; ++k) {
            float o[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                if (((MSH + v) & 1) == 0) o[v] = k == S - 1 ? sg[v] : ((k & 1) ? pr[v][k >> 1].y : pr[v][k >> 1].x);
                else o[v] = k == 0 ? sg[v] : (((k - 1) & 1) ? pr[v][(k - 1) >> 1].y : pr[v][(k - 1) >> 1].x);
            }
            ca_st_stream4(reinterpret_cast<float4*>(rp + k * a.rl.plane), make_float4(o[0], o[1], o[2], o[3]));
            if (lpad) ca_st_stream4(reinterpret_cast<float4*>(rp + k * a.rl.plane - 4), zero4);
            if (rpad) ca_st_stream4(reinterpret_cast<float4*>(rp + k * a.rl.plane + 4), zero4);
        }
    }
}

// ---- the two box filters, fused ------------------------------------------------------------------------------------
struct Box12Args {
    const float* raw;
    float* ssd;
    RawLayout rl;
    int w, d, S, chunk_start;
    int r0, r1;             // raw planes present in the chunk buffer: global H range [r0, r1)
    int b0, b1;             // first-box planes the outputs need: [b0, b1) = [xb-1, xe+1) clipped to the volume
    int xb, xe;             // output planes [xb, xe) of the (L, xe-xb, w, d) slab
    int nq, RP, Ty;         // quads per row, row pairs per CTA (first box), output rows per CTA (second box) = 2*(RP-1)
};

// Per-thread state of the march along H.  Rows that do not exist (outside the buffer, padding threads) point at an
// all-zero row, so the inner loop carries no validity predicates.
struct BoxThread {
    const float* rawp[4];   // the thread's 4 raw rows (y0-2 .. y0+1 of the first box's tile) at plane X
    const float* p1r[4];    // the same 4 row slots of the first box's plane in shared memory (buffer 0)
    float* p1w[2];          // where the thread's two first-box rows go (buffer 0); nullptr: row is outside the volume
    float* out[2];          // the thread's two cost-volume rows at plane X-2; nullptr: not an output row
    int p1buf;              // floats between the two shared-memory plane buffers
    int64_t xPitch, outPitch;
    BoxLane bl;
    bool zok[4];            // z0+i < d
};

template <int R, bool DALIGNED>
static __device__ __forceinline__ void box12_step(int X, const Box12Args& a, BoxThread& t, float (&s1)[3][2][4],
                                                  float (&s2)[3][2][4], float4 (&cur)[4]) {
    constexpr int Cs = (R + 1) % 3;
    if (X > a.xe + 1) return;
    const int boff = (X & 1) * t.p1buf;
    // ---- first box: raw plane X in, box1 plane X-1 out (shared memory) ----
    if (X < a.r1) {
        float hl[4], hr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) box_halo(cur[j], t.rawp[j], t.bl, hl[j], hr[j]);
        box_add_plane<R>(s1, cur, hl, hr);
#pragma unroll
        for (int j = 0; j < 4; ++j) t.rawp[j] += t.xPitch;
    }
    const bool p1ok = X - 1 >= a.b0 && X - 1 < a.b1;
    if (p1ok) {
#pragma unroll
        for (int o = 0; o < 2; ++o)
            if (t.p1w[o]) {
                float4 r;
                r.x = ca_div27(s1[Cs][o][0]); r.y = ca_div27(s1[Cs][o][1]); r.z = ca_div27(s1[Cs][o][2]); r.w = ca_div27(s1[Cs][o][3]);
                if (!DALIGNED) { r.y = t.zok[1] ? r.y : 0.f; r.z = t.zok[2] ? r.z : 0.f; r.w = t.zok[3] ? r.w : 0.f; }
                *reinterpret_cast<float4*>(t.p1w[o] + boff) = r;
            }
    }
    __syncthreads();
    // ---- next raw plane: issued now, consumed after the second box's arithmetic ----
    if (X + 1 < a.r1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) cur[j] = __ldg(reinterpret_cast<const float4*>(t.rawp[j]));
    }
    // ---- second box: box1 plane P = X-1 in, cost-volume plane P-1 out ----
    if (p1ok) {
        float4 v[4];
        float hl[4], hr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float* quad = t.p1r[j] + boff;
            v[j] = *reinterpret_cast<const float4*>(quad);
            box_halo(v[j], quad, t.bl, hl[j], hr[j]);
        }
        box_add_plane<R>(s2, v, hl, hr);
    } else {
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int k = 0; k < 4; ++k) s2[R][o][k] = 0.f;      // a plane outside the volume: the fresh sum starts at 0
    }
    if (X - 2 >= a.xb) {
#pragma unroll
        for (int o = 0; o < 2; ++o)
            if (t.out[o]) {
                const float r0 = ca_div27(s2[Cs][o][0]), r1 = ca_div27(s2[Cs][o][1]), r2 = ca_div27(s2[Cs][o][2]),
                            r3 = ca_div27(s2[Cs][o][3]);
                if (DALIGNED) {
                    ca_st_stream4(reinterpret_cast<float4*>(t.out[o]), make_float4(r0, r1, r2, r3));
                } else {
                    if (t.zok[0]) t.out[o][0] = r0;
                    if (t.zok[1]) t.out[o][1] = r1;
                    if (t.zok[2]) t.out[o][2] = r2;
                    if (t.zok[3]) t.out[o][3] = r3;
                }
                t.out[o] += t.outPitch;
            }
    }
}


// grid = (row tiles, planes of the chunk); block = nq*RP threads rounded up to whole warps, z-fastest.
// Shared memory: two plane buffers of (2*RP rows + one all-zero row) x 4*nq floats.
template <bool DALIGNED>
__global__ void __launch_bounds__(512, 1) k_box12(Box12Args a) {
    CA_DYN_SMEM(float, p1);
    const int tid = (int)threadIdx.x;
    const int q = tid % a.nq, rp = tid / a.nq;
    const int lane = tid & 31;
    const int z0 = 4 * q;
    const bool active1 = rp < a.RP;                 // threads beyond nq*RP only keep the warps whole
    const int ty0 = (int)blockIdx.x * a.Ty;
    const int pl = (int)blockIdx.y;
    const int g = pl / a.S, kD = pl - g * a.S;
    const int64_t label = (int64_t)kD * a.S * a.S + a.chunk_start + g;
    const int p1pitch = 4 * a.nq;
    BoxThread t;
    t.p1buf = (2 * a.RP + 1) * p1pitch;
    t.xPitch = a.rl.xPitch;
    t.outPitch = (int64_t)a.w * a.d;
    t.bl.first = q == 0;
    t.bl.last = q == a.nq - 1;
    t.bl.fb_l = !t.bl.first && lane == 0;
    t.bl.fb_r = !t.bl.last && lane == 31;
#pragma unroll
    for (int i = 0; i < 4; ++i) t.zok[i] = z0 + i < a.d;
    float* zero_row = p1 + 2 * a.RP * p1pitch;
    for (int i = tid; i < p1pitch; i += (int)blockDim.x) zero_row[i] = zero_row[t.p1buf + i] = 0.f;
    // raw rows y = ty0 - 2 + 2*rp + j  <->  buffer rows ty0 + 2*rp + j; buffer row 0 is all zero
    const float* plane = a.raw + (int64_t)pl * a.rl.plane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int br = ty0 + 2 * rp + j;
        t.rawp[j] = plane + (int64_t)((active1 && br < a.w + 4) ? br : 0) * a.rl.rowPitch + 4 + z0;
        const int e = 2 * rp + j;
        t.p1r[j] = (active1 && e < 2 * a.RP) ? p1 + e * p1pitch + z0 : zero_row + z0;
    }
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const int e = 2 * rp + o;
        const int y1 = ty0 - 1 + e;                 // first-box row
        const bool in1 = active1 && y1 >= 0 && y1 < a.w;
        t.p1w[o] = in1 ? p1 + e * p1pitch + z0 : nullptr;
        if (active1 && !in1) {                      // rows outside the volume stay zero in both buffers
            *reinterpret_cast<float4*>(p1 + e * p1pitch + z0) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(p1 + t.p1buf + e * p1pitch + z0) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const int y2 = ty0 + e;                     // second-box (output) row
        const bool in2 = rp < a.RP - 1 && y2 < a.w;
        t.out[o] = in2 ? a.ssd + (label * (a.xe - a.xb) * a.w + y2) * a.d + z0 : nullptr;
    }
    float4 cur[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cur[j] = __ldg(reinterpret_cast<const float4*>(t.rawp[j]));
    __syncthreads();
    // march over the global planes r0 .. xe+1 (the second box lags one plane behind the first)
    float s1[3][2][4], s2[3][2][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int k = 0; k < 4; ++k) s1[i][o][k] = s2[i][o][k] = 0.f;
    for (int X = a.r0; X <= a.xe + 1; X += 3) {
        box12_step<0, DALIGNED>(X, a, t, s1, s2, cur);
        box12_step<1, DALIGNED>(X + 1, a, t, s1, s2, cur);
        box12_step<2, DALIGNED>(X + 2, a, t, s1, s2, cur);
    }
}

// ---- the two box filters, FAST variant: separable sums, raw planes staged by the TMA engine ---------------------------
// box3 o box3 with the truncation of the intermediate to the volume is separable per axis, so each box is a 3-tap sum
// along D, along W and along H (the in-plane sums of the two previous planes): 8 additions and one multiplication by
// RN(1/27) per value and box instead of 26 additions and an exact division.  The result differs from the reference's
// ordered sum by a few ulp (<= 30 roundings against <= 68, see coupled.cu "certified selection"), which is why this mode
// never decides an argmin on its own.
// Same tiling and march as k_box12, but the kernel is bound by memory latency, not by issue slots, so the raw planes do
// not go through registers: the tile's rows of plane X+NST (one contiguous run of the chunk buffer, zero halo rows and
// pad columns included) are requested with ONE bulk asynchronous copy (cp.async.bulk + mbarrier) into a ring of NST
// shared-memory stages while plane X is being filtered, and every tap -- the +-1 neighbours along D included -- is a
// shared-memory read of a zero-padded row: no shuffles, no boundary selects, no address arithmetic per row.
#ifndef CA_BOXF_NST
#define CA_BOXF_NST 4
#endif
struct BoxFastArgs {
    const float* raw;
    float* ssd;
    RawLayout rl;
    int w, d, S, chunk_start;
    int r0, r1, b0, b1, xb, xe;
    int nq, RP, Ty;
};

// sums of a 3x3 window (rows j..j+2, columns -1..+1) for the thread's two output rows x 4 columns; p = the thread's
// first row at column z0, pitch in floats, columns z0-1 and z0+4 readable (zero pads)
static __device__ __forceinline__ void boxf_inplane(const float* p, int pitch, float (&out)[2][4]) {
    float ds[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float* r = p + j * pitch;
        const float4 v = *reinterpret_cast<const float4*>(r);
        const float hl = r[-1], hr = r[4];
        ds[j][0] = (hl + v.x) + v.y;
        ds[j][1] = (v.x + v.y) + v.z;
        ds[j][2] = (v.y + v.z) + v.w;
        ds[j][3] = (v.z + v.w) + hr;
    }
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int k = 0; k < 4; ++k) out[o][k] = (ds[o][k] + ds[o + 1][k]) + ds[o + 2][k];
}

// grid = (row tiles, planes of the chunk); block = nq*RP threads rounded up to whole warps.
// shared memory: NST raw stages of (2*RP+2) rows, then two first-box planes of 2*RP (+2 zero) rows; row pitch dF+8, data at column 4
template <bool DALIGNED>
__global__ void __launch_bounds__(512, 1) k_box12_fast(BoxFastArgs a) {
    CA_DYN_SMEM(float, sm);
    __shared__ unsigned long long full[CA_BOXF_NST];
    const int tid = (int)threadIdx.x, nthr = (int)blockDim.x;
    const int q = tid % a.nq, rp = tid / a.nq;
    const int z0 = 4 * q;
    const int pitch = a.rl.rowPitch;
    const int RT = 2 * a.RP + 2;                         // raw rows of the tile
    const int stage_f = RT * pitch, p1_f = RT * pitch;       // (the last row pair's threads read two rows past the 2*RP first-box rows: zeros)
    float* sraw = sm;
    float* sp1 = sm + CA_BOXF_NST * stage_f;
    const int ty0 = (int)blockIdx.x * a.Ty;
    const int pl = (int)blockIdx.y;
    const int g = pl / a.S, kD = pl - g * a.S;
    const int64_t label = (int64_t)kD * a.S * a.S + a.chunk_start + g;
    // tile rows [ty0, ty0+RT) of the chunk buffer (row r <-> y = r-2); only the interior rows [2, w+2) are copied: the halo
    // rows are not written by the staged raw kernel and stay zero in the stages
    const int rfirst = ca_max(ty0, 2), rlast = ca_min(ty0 + RT, a.w + 2);
    const unsigned copy_bytes = (unsigned)((rlast - rfirst) * pitch) * 4;
    const int stage_off = (rfirst - ty0) * pitch;
    const float* src0 = a.raw + (int64_t)pl * a.rl.plane + (int64_t)rfirst * pitch;
    // everything that is never copied or written stays zero: missing raw rows, first-box rows outside the volume, pads
    for (int i = tid; i < CA_BOXF_NST * stage_f + 2 * p1_f; i += nthr) sm[i] = 0.f;
    if (tid == 0)
        for (int s = 0; s < CA_BOXF_NST; ++s) ca_mbar_init(&full[s], 1);
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < CA_BOXF_NST && a.r0 + s < a.r1; ++s) {
            ca_mbar_expect_tx(&full[s], copy_bytes);
            ca_bulk_g2s(sraw + s * stage_f + stage_off, src0 + (int64_t)s * a.rl.xPitch, copy_bytes, &full[s]);
        }
    const bool active1 = rp < a.RP;
    bool w1[2], w2[2];                                   // the thread's first-box rows / output rows exist
    float* out[2];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
        const int y1 = ty0 - 1 + 2 * rp + o, y2 = ty0 + 2 * rp + o;
        w1[o] = active1 && y1 >= 0 && y1 < a.w;
        w2[o] = rp < a.RP - 1 && y2 < a.w;
        out[o] = a.ssd + (label * (a.xe - a.xb) * a.w + (w2[o] ? y2 : 0)) * a.d + z0;
    }
    const int roff = (active1 ? 2 * rp : 0) * pitch + 4 + z0;      // the thread's first row in a stage / first-box plane
    const int64_t outPitch = (int64_t)a.w * a.d;
    const float c27 = 1.0f / 27.0f;
    float h1[2][2][4], h2[2][2][4];                      // in-plane sums of the previous plane [0] and the one before [1]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int k = 0; k < 4; ++k) h1[i][o][k] = h2[i][o][k] = 0.f;
    for (int X = a.r0; X <= a.xe + 1; ++X) {
        const int it = X - a.r0, st = it % CA_BOXF_NST;
        float in[2][4];
        // ---- first box: raw plane X in, box1 plane X-1 out ----
        if (X < a.r1) {
            ca_mbar_wait(&full[st], (unsigned)(it / CA_BOXF_NST) & 1u);
            boxf_inplane(sraw + st * stage_f + roff, pitch, in);
        } else {
#pragma unroll
            for (int o = 0; o < 2; ++o)
#pragma unroll
                for (int k = 0; k < 4; ++k) in[o][k] = 0.f;
        }
        const bool p1ok = X - 1 >= a.b0 && X - 1 < a.b1;
        float* p1 = sp1 + (X & 1) * p1_f;
        if (p1ok) {
#pragma unroll
            for (int o = 0; o < 2; ++o)
                if (w1[o]) {
                    float4 r;
                    r.x = ((h1[1][o][0] + h1[0][o][0]) + in[o][0]) * c27;
                    r.y = ((h1[1][o][1] + h1[0][o][1]) + in[o][1]) * c27;
                    r.z = ((h1[1][o][2] + h1[0][o][2]) + in[o][2]) * c27;
                    r.w = ((h1[1][o][3] + h1[0][o][3]) + in[o][3]) * c27;
                    if (!DALIGNED) { r.y = z0 + 1 < a.d ? r.y : 0.f; r.z = z0 + 2 < a.d ? r.z : 0.f; r.w = z0 + 3 < a.d ? r.w : 0.f; }
                    *reinterpret_cast<float4*>(p1 + roff + o * pitch) = r;
                }
        }
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int k = 0; k < 4; ++k) { h1[1][o][k] = h1[0][o][k]; h1[0][o][k] = in[o][k]; }
        __syncthreads();
        // every thread is done with stage st: refill it with plane X + NST
        if (tid == 0 && X + CA_BOXF_NST < a.r1) {
            ca_mbar_expect_tx(&full[st], copy_bytes);
            ca_bulk_g2s(sraw + st * stage_f + stage_off, src0 + (int64_t)(it + CA_BOXF_NST) * a.rl.xPitch, copy_bytes, &full[st]);
        }
        // ---- second box: box1 plane X-1 in, cost-volume plane X-2 out ----
        if (p1ok) {
            boxf_inplane(p1 + roff, pitch, in);
        } else {
#pragma unroll
            for (int o = 0; o < 2; ++o)
#pragma unroll
                for (int k = 0; k < 4; ++k) in[o][k] = 0.f;     // a plane outside the volume
        }
        if (X - 2 >= a.xb) {
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                if (w2[o]) {
                    const float r0 = ((h2[1][o][0] + h2[0][o][0]) + in[o][0]) * c27, r1 = ((h2[1][o][1] + h2[0][o][1]) + in[o][1]) * c27,
                                r2 = ((h2[1][o][2] + h2[0][o][2]) + in[o][2]) * c27, r3 = ((h2[1][o][3] + h2[0][o][3]) + in[o][3]) * c27;
                    if (DALIGNED) {
                        ca_st_stream4(reinterpret_cast<float4*>(out[o]), make_float4(r0, r1, r2, r3));
                    } else {
                        if (z0 + 0 < a.d) out[o][0] = r0;
                        if (z0 + 1 < a.d) out[o][1] = r1;
                        if (z0 + 2 < a.d) out[o][2] = r2;
                        if (z0 + 3 < a.d) out[o][3] = r3;
                    }
                }
                out[o] += outPitch;
            }
        }
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int k = 0; k < 4; ++k) { h2[1][o][k] = h2[0][o][k]; h2[0][o][k] = in[o][k]; }
    }
}

namespace {
struct CorrPlan {
    int S, dF, dP, G;
    int nq, RP, Ty, ntiles, nthreads;
    int r0, r1, b0, b1;
    RawLayout rl;
    size_t off_fix, off_mov, off_raw, off_keys, total, chunk_bytes, box_smem, boxf_smem;
};

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// G = number of (kH,kW) pairs per chunk; G<=0 -> default (raw chunk buffer of up to ~3 GB: a handful of launches).
// [xb, xe) = the H planes of the cost volume to produce (a slab of a spatially split pair; the whole volume: 0, h).
CorrPlan make_plan(int C, int h, int w, int d, int hw, int xb, int xe, int G, bool fast = false) {
    CorrPlan p;
    p.S = 2 * hw + 1;
    p.dF = ca_roundup4(d);
    p.dP = p.dF + ca_roundup4(4 + 2 * hw) - 4;
    p.r0 = ca_max(0, xb - 2); p.r1 = ca_min(h, xe + 2);
    p.b0 = ca_max(0, xb - 1); p.b1 = ca_min(h, xe + 1);
    p.rl = make_raw_layout(p.r1 - p.r0, w, p.dF);
    // box12 tiling: whole rows per CTA (nq quads), RP row pairs of the first box -> Ty = 2*(RP-1) output rows
    p.nq = p.dF / 4;
    static int env_bt = -1;      // tuning knob: CA_BOX_THREADS = largest CTA of the box kernel (default 512)
    if (env_bt < 0) { const char* e = getenv("CA_BOX_THREADS"); env_bt = e ? atoi(e) : 0; }
    // the fast kernel is bound by shared-memory latency and the per-plane barrier: two or three smaller CTAs per SM
    // overlap better than one of 480 threads (measured at config 2: 8.0 vs 8.7 ms per direction)
    const int bt = (env_bt >= 64 && env_bt <= 512) ? env_bt : (fast ? 320 : 512);
    int rp_max = bt / p.nq;
    if (rp_max > (w + 3) / 2 + 1) rp_max = (w + 3) / 2 + 1;     // one tile covers all rows
    if (rp_max < 2) rp_max = 2;
    const int ty_max = 2 * (rp_max - 1);
    p.ntiles = (w + ty_max - 1) / ty_max;
    p.Ty = (((w + p.ntiles - 1) / p.ntiles) + 1) & ~1;          // balanced tiles, even
    p.RP = p.Ty / 2 + 1;
    p.nthreads = (p.nq * p.RP + 31) & ~31;
    p.box_smem = (size_t)2 * (2 * p.RP + 1) * (4 * p.nq) * sizeof(float);
    p.boxf_smem = (size_t)(CA_BOXF_NST + 2) * (2 * p.RP + 2) * (p.dF + 8) * sizeof(float);
    const size_t plane = (size_t)h * w * p.dF * 4;
    const size_t pair_bytes = (size_t)p.S * p.rl.plane * 4;
    if (G <= 0) {
        static int env_g = -1;   // tuning knob: CA_CORR_CHUNK = (kH,kW) pairs per chunk
        if (env_g < 0) { const char* e = getenv("CA_CORR_CHUNK"); env_g = e ? atoi(e) : 0; }
        G = env_g > 0 ? env_g : (int)(((size_t)3 << 30) / pair_bytes);
        if (G < 1) G = 1;
        if (G > p.S) G -= G % p.S;      // whole kW groups: the staged raw kernel shares one moving tile among them
    }
    if (G > p.S * p.S) G = p.S * p.S;
    p.G = G;
    size_t o = 0;
    p.off_fix = o; o = align256(o + (size_t)C * plane);
    p.off_mov = o; o = align256(o + (size_t)C * (h + 2 * hw) * (w + 2 * hw) * p.dP * 4);
    p.chunk_bytes = (size_t)G * pair_bytes;
    p.off_raw = o; o = align256(o + p.chunk_bytes);
    p.off_keys = o; o = align256(o + (size_t)(xe - xb) * w * d * 8);
    p.total = o;
    return p;
}

static bool use_staged() {       // CA_RAW_STAGED=0: A/B switch back to the direct-from-L2 raw kernel
    static int v = -1;
    if (v < 0) { const char* e = getenv("CA_RAW_STAGED"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v == 1;
}
static int g_raw_minb = 0;   // tuning knob (CA_RAW_MINB env, read once): min resident blocks per SM of k_raw_ssd

template <int HW, bool FAST>
void launch_raw_t(const float* fixp, const float* movp, int C, int h, int w, int d, int chunk_start, int g, const CorrPlan& p,
                  float* raw, cudaStream_t st) {
    const TileGeom t = make_tile(p.dF, 128);
    const dim3 grid((unsigned)g, (unsigned)(t.nzb * ((w + t.by - 1) / t.by)), (unsigned)(p.r1 - p.r0)), block(t.bx, t.by);
    if (g_raw_minb == 0) {
        const char* e = getenv("CA_RAW_MINB");
        g_raw_minb = e ? atoi(e) : 6;
        if (g_raw_minb != 4 && g_raw_minb != 5 && g_raw_minb != 8) g_raw_minb = 6;
    }
    if (g_raw_minb == 5) { auto k = k_raw_ssd<HW, 5, FAST>; CA_LAUNCH(k, grid, block, 0, st, fixp, movp, C, h, w, d, p.dF, p.dP, chunk_start, t.zq, t.nzb, p.rl, p.r0, raw); }
    else if (g_raw_minb == 4) { auto k = k_raw_ssd<HW, 4, FAST>; CA_LAUNCH(k, grid, block, 0, st, fixp, movp, C, h, w, d, p.dF, p.dP, chunk_start, t.zq, t.nzb, p.rl, p.r0, raw); }
    else if (g_raw_minb == 8) { auto k = k_raw_ssd<HW, 8, FAST>; CA_LAUNCH(k, grid, block, 0, st, fixp, movp, C, h, w, d, p.dF, p.dP, chunk_start, t.zq, t.nzb, p.rl, p.r0, raw); }
    else { auto k = k_raw_ssd<HW, 6, FAST>; CA_LAUNCH(k, grid, block, 0, st, fixp, movp, C, h, w, d, p.dF, p.dP, chunk_start, t.zq, t.nzb, p.rl, p.r0, raw); }
}
template <int HW>
void launch_raw(bool fast, const float* fixp, const float* movp, int C, int h, int w, int d, int chunk_start, int g, const CorrPlan& p,
                float* raw, cudaStream_t st) {
    if (fast) launch_raw_t<HW, true>(fixp, movp, C, h, w, d, chunk_start, g, p, raw, st);
    else launch_raw_t<HW, false>(fixp, movp, C, h, w, d, chunk_start, g, p, raw, st);
}

// smem / thread limits of the staged kernel for NW kW groups per chunk
static bool staged_fits(const CorrPlan& p, int C, int NW) {
    const size_t smem = ((size_t)C * CA_RAW_TY * p.dF + (size_t)C * (CA_RAW_TY + NW - 1) * p.dP) * sizeof(float);
    return smem <= 110 * 1024 && (p.dF / 4) * CA_RAW_TY <= 256;
}
// true when every chunk of the plan is a set of whole kW groups the staged kernel can take
static bool all_chunks_staged(const CorrPlan& p, int C) {
    return use_staged() && p.G % p.S == 0 && staged_fits(p, C, p.G / p.S);
}

// staged raw kernel for a chunk of whole kW groups (chunk_start = kW0 * S, g = NW * S); false: shape not supported
template <int HW>
bool launch_raw_staged(bool fast, const float* fixp, const float* movp, int C, int h, int w, int d, int kW0, int NW, const CorrPlan& p,
                       float* raw, cudaStream_t st, int* rc) {
    constexpr int S = 2 * HW + 1;
    const size_t smem = ((size_t)C * CA_RAW_TY * p.dF + (size_t)C * (CA_RAW_TY + NW - 1) * p.dP) * sizeof(float);
    const int nq = p.dF / 4;
    if (!staged_fits(p, C, NW)) return false;                        // >= two CTAs per SM, <= 256 threads
    RawStagedArgs a;
    a.fixp = fixp; a.movp = movp; a.raw = raw; a.rl = p.rl;
    a.C = C; a.h = h; a.w = w; a.d = d; a.dF = p.dF; a.dP = p.dP; a.r0 = p.r0; a.kW0 = kW0; a.NW = NW;
    const dim3 grid((unsigned)S, (unsigned)((w + CA_RAW_TY - 1) / CA_RAW_TY), (unsigned)(p.r1 - p.r0)), block((unsigned)nq, CA_RAW_TY);
    *rc = [&]() -> int {
        if (fast) { auto k = k_raw_ssd_staged<HW, true>; CA_SMEM_OPTIN(k, 110 * 1024); CA_LAUNCH(k, grid, block, smem, st, a); }
        else { auto k = k_raw_ssd_staged<HW, false>; CA_SMEM_OPTIN(k, 110 * 1024); CA_LAUNCH(k, grid, block, smem, st, a); }
        return CA_OK;
    }();
    return true;
}

// geometry of the tensor-map raw kernel for NW kW groups per chunk; false: shape outside its range (d a multiple of 4,
// boxes of at most 256 elements per dimension, two CTAs per SM)
static bool tmap_geometry(int C, int d, int hw, int NW, int* fix_f, int* mz, size_t* smem) {
    const int msh = (4 - hw % 4) % 4, nm = (msh + 4 + 2 * hw + 3) & ~3;
    *mz = d - 4 + nm;
    *fix_f = (C * CA_RAW_TY * d + 31) & ~31;
    *smem = ((size_t)*fix_f + (size_t)C * (CA_RAW_TY + NW - 1) * *mz) * sizeof(float) + 128;
    return (d & 3) == 0 && d <= 240 && *mz <= 256 && C <= 256 && *smem <= 110 * 1024 && (d / 4) * CA_RAW_TY <= 256;
}
static bool use_tmap() {         // CA_RAW_TMAP=0: A/B switch back to the bulk-copy staged kernel on the zero-padded copy
    static int v = -1;
    if (v < 0) { const char* e = getenv("CA_RAW_TMAP"); v = (e && atoi(e) == 0) ? 0 : 1; }
    return v == 1;
}
// true when every chunk of the plan goes through the tensor-map kernel (certified mode, unpadded features in place)
static bool all_chunks_tmap(const CorrPlan& p, int C, int d, int hw, const float* fix, const float* mov) {
    int ff, mz; size_t sm;
    return use_tmap() && use_staged() && p.G % p.S == 0 && ((((uintptr_t)fix) | ((uintptr_t)mov)) & 15) == 0 &&
           tmap_geometry(C, d, hw, p.G / p.S, &ff, &mz, &sm);
}

template <int HW>
int launch_raw_tmap(const float* fix, const float* mov, int C, int h, int w, int d, int kW0, int NW, const CorrPlan& p, float* raw,
                    cudaStream_t st) {
    constexpr int S = 2 * HW + 1;
    RawTmapArgs a;
    size_t smem;
    if (!tmap_geometry(C, d, HW, NW, &a.fix_f, &a.mz, &smem)) return CA_ERR_UNSUPPORTED;
    a.raw = raw; a.rl = p.rl; a.C = C; a.w = w; a.d = d; a.r0 = p.r0; a.kW0 = kW0; a.NW = NW;
    CaTensorMap tmF, tmM;
    int rc = ca_tmap_make(&tmF, fix, d, w, h, C, d, d, CA_RAW_TY, C);
    if (rc != CA_OK) return rc;
    rc = ca_tmap_make(&tmM, mov, d, w, h, C, d, a.mz, CA_RAW_TY + NW - 1, C);
    if (rc != CA_OK) return rc;
    const dim3 grid((unsigned)S, (unsigned)((w + CA_RAW_TY - 1) / CA_RAW_TY), (unsigned)(p.r1 - p.r0)), block((unsigned)(d / 4), CA_RAW_TY);
    auto k = k_raw_ssd_tmap<HW>;
    CA_SMEM_OPTIN(k, 110 * 1024);
    CA_LAUNCH(k, grid, block, smem, st, tmF, tmM, a);
    return CA_OK;
}
}  // namespace

extern "C" size_t ca_correlate_slab_workspace_bytes(int C, int h, int w, int d, int disp_hw, int x_begin, int x_end) {
    if (C <= 0 || h <= 0 || w <= 0 || d <= 0 || disp_hw < 1 || x_begin < 0 || x_end > h || x_begin >= x_end) return 0;
    return make_plan(C, h, w, d, disp_hw, x_begin, x_end, 0).total;
}
extern "C" size_t ca_correlate_workspace_bytes(int C, int h, int w, int d, int disp_hw) {
    return ca_correlate_slab_workspace_bytes(C, h, w, d, disp_hw, 0, h);
}
extern "C" size_t ca_correlate_min_workspace_bytes(int C, int h, int w, int d, int disp_hw) {
    if (C <= 0 || h <= 0 || w <= 0 || d <= 0 || disp_hw < 1) return 0;
    return make_plan(C, h, w, d, disp_hw, 0, h, 1).total;
}

static int correlate_impl(bool fast, const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, int x_begin,
                          int x_end, float* ssd, int64_t* argmin, void* workspace, size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(fix && mov && ssd && workspace, "correlate: null pointer");
    CA_CHECK_ARG(!(fast && argmin), "correlate (fast): the argmin of approximate costs is not the reference's; use ca_coupled_convex_certified");
    CA_CHECK_ARG(C > 0 && h > 0 && w > 0 && d > 0, "correlate: bad dims");
    CA_CHECK_ARG(x_begin >= 0 && x_end <= h && x_begin < x_end, "correlate: bad slab [%d, %d) of %d planes", x_begin, x_end, h);
    if (disp_hw < 1 || disp_hw > 8) {
        ca_set_error("correlate: disp_hw=%d unsupported (1..8)", disp_hw);
        return CA_ERR_UNSUPPORTED;
    }
    if (d > 1024) {
        ca_set_error("correlate: coarse depth %d unsupported (<= 1024)", d);
        return CA_ERR_UNSUPPORTED;
    }
    CA_CHECK_ARG((int64_t)h * (w + 4) * (ca_roundup4(d) + 8) < (1ll << 31) && h < 65536, "correlate: coarse grid too large");
    // largest chunk that fits the caller's workspace (>= 1 pair), capped at the default plan
    CorrPlan p = make_plan(C, h, w, d, disp_hw, x_begin, x_end, 0, fast);
    if (p.total > workspace_bytes) {
        const CorrPlan p1 = make_plan(C, h, w, d, disp_hw, x_begin, x_end, 1);
        if (p1.total > workspace_bytes) {
            ca_set_error("correlate: workspace too small (%zu < %zu)", workspace_bytes, p1.total);
            return CA_ERR_WORKSPACE;
        }
        const size_t pair_bytes = p1.chunk_bytes;
        int G = 1 + (int)((workspace_bytes - p1.total) / pair_bytes);
        p = make_plan(C, h, w, d, disp_hw, x_begin, x_end, G, fast);
        while (p.total > workspace_bytes && p.G > 1) p = make_plan(C, h, w, d, disp_hw, x_begin, x_end, p.G - 1, fast);
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    float* fixp = (float*)(ws + p.off_fix);
    float* movp = (float*)(ws + p.off_mov);
    float* raw = (float*)(ws + p.off_raw);
    const int S = p.S, npairs = S * S;
    if (fast) {
        // One fused kernel (correlate_fused.cu) straight from the unpadded feature volumes: tensor-map TMA needs rows of a
        // multiple of 4 floats on 16-byte aligned bases, anything else goes through a pitched copy first.
        const float *ff = fix, *mm = mov;
        const bool direct = p.dF == d && ((((uintptr_t)fix) | ((uintptr_t)mov)) & 15) == 0;
        int rc = CA_OK;
        if (!direct) {
            const int64_t tf = (int64_t)C * h * w * p.dF;
            const unsigned nb = (unsigned)ca_min((int)ca_divup(tf, 256), 148 * 16);
            CA_LAUNCH(k_pad_features, dim3(nb), dim3(256), 0, st, fix, C, h, w, d, 0, p.dF, fixp);
            CA_LAUNCH(k_pad_features, dim3(nb), dim3(256), 0, st, mov, C, h, w, d, 0, p.dF, movp);
            CA_CHECK_LAUNCH("correlate/pitch");
            ff = fixp; mm = movp;
        }
        rc = ca_launch_corr_fused(ff, mm, C, h, w, d, p.dF, disp_hw, x_begin, x_end, ssd, st);
        if (rc != CA_ERR_UNSUPPORTED) return rc;
        // default: one launch of producer / consumer CTAs, raw planes through a ring in L2 (the chunk buffer's space)
        if (direct) {
            rc = ca_launch_corr_pipe(fix, mov, C, h, w, d, disp_hw, x_begin, x_end, ssd, ws + p.off_raw, workspace_bytes - p.off_raw, st);
            if (rc != CA_ERR_UNSUPPORTED) return rc;
        }
    }
    // certified mode, rows of a multiple of 4 voxels: the raw kernel stages its tiles by tensor-map TMA straight from the
    // caller's feature volumes (out-of-bounds zero fill = F.pad): no padded copy, no halo zeroing
    const bool tmap = fast && all_chunks_tmap(p, C, d, disp_hw, fix, mov);
    if (!tmap) {
        const int64_t tf = (int64_t)C * h * w * p.dF, tm = (int64_t)C * (h + 2 * disp_hw) * (w + 2 * disp_hw) * p.dP;
        // the fixed features are used in place when their rows already have the kernels' pitch and alignment
        if (p.dF == d && (((uintptr_t)fix) & 15) == 0) fixp = const_cast<float*>(fix);
        else CA_LAUNCH(k_pad_features, dim3((unsigned)ca_min((int)ca_divup(tf, 256), 148 * 16)), dim3(256), 0, st, fix, C, h, w, d, 0, p.dF, fixp);
        CA_LAUNCH(k_pad_features, dim3((unsigned)ca_min((int)ca_divup(tm, 256), 148 * 16)), dim3(256), 0, st, mov, C, h, w, d, disp_hw, p.dP, movp);
        CA_CHECK_LAUNCH("correlate/pad");
        // zero the halo rows / pad columns of the chunk buffer once; the raw kernels only ever write interiors
        // (the staged raw kernel writes the halo of its own planes; only the direct kernel needs it done up front)
        if (!fast || !all_chunks_staged(p, C)) {
            const int64_t nslices = (int64_t)p.G * p.S * (p.r1 - p.r0);
            CA_LAUNCH(k_zero_raw_halo, dim3((unsigned)(nslices < 148 * 32 ? nslices : 148 * 32)), dim3(128), 0, st, raw, nslices, w, p.rl.rowPitch);
            CA_CHECK_LAUNCH("correlate/zero_halo");
        }
    }
    if (p.box_smem > 160 * 1024 || (fast && p.boxf_smem > 200 * 1024)) {
        ca_set_error("correlate: row tile needs %zu bytes of shared memory", p.box_smem);
        return CA_ERR_UNSUPPORTED;
    }
    // (balancing the chunks, 4+3+3+3 kW groups instead of 4+4+4+1, was measured and is slower: 7.99 vs 7.89 ms per direction)
    for (int cs = 0; cs < npairs; cs += p.G) {
        const int g = ca_min(p.G, npairs - cs);
        void* tm = ca_timing_begin(CA_T_RAW, st);
        bool staged = false;
        int src = CA_OK;
        if (tmap) {
            switch (disp_hw) {
                case 1: src = launch_raw_tmap<1>(fix, mov, C, h, w, d, cs / S, g / S, p, raw, st); break;
                case 2: src = launch_raw_tmap<2>(fix, mov, C, h, w, d, cs / S, g / S, p, raw, st); break;
                case 3: src = launch_raw_tmap<3>(fix, mov, C, h, w, d, cs / S, g / S, p, raw, st); break;
                case 4: src = launch_raw_tmap<4>(fix, mov, C, h, w, d, cs / S, g / S, p, raw, st); break;
                case 5: src = launch_raw_tmap<5>(fix, mov, C, h, w, d, cs / S, g / S, p, raw, st); break;
                case 6: src = launch_raw_tmap<6>(fix, mov, C, h, w, d, cs / S, g / S, p, raw, st); break;
                case 7: src = launch_raw_tmap<7>(fix, mov, C, h, w, d, cs / S, g / S, p, raw, st); break;
                default: src = launch_raw_tmap<8>(fix, mov, C, h, w, d, cs / S, g / S, p, raw, st); break;
            }
            staged = true;
        } else if (use_staged() && cs % S == 0 && g % S == 0) {
            switch (disp_hw) {
                case 1: staged = launch_raw_staged<1>(fast, fixp, movp, C, h, w, d, cs / S, g / S, p, raw, st, &src); break;
                case 2: staged = launch_raw_staged<2>(fast, fixp, movp, C, h, w, d, cs / S, g / S, p, raw, st, &src); break;
                case 3: staged = launch_raw_staged<3>(fast, fixp, movp, C, h, w, d, cs / S, g / S, p, raw, st, &src); break;
                case 4: staged = launch_raw_staged<4>(fast, fixp, movp, C, h, w, d, cs / S, g / S, p, raw, st, &src); break;
                case 5: staged = launch_raw_staged<5>(fast, fixp, movp, C, h, w, d, cs / S, g / S, p, raw, st, &src); break;
                case 6: staged = launch_raw_staged<6>(fast, fixp, movp, C, h, w, d, cs / S, g / S, p, raw, st, &src); break;
                case 7: staged = launch_raw_staged<7>(fast, fixp, movp, C, h, w, d, cs / S, g / S, p, raw, st, &src); break;
                default: staged = launch_raw_staged<8>(fast, fixp, movp, C, h, w, d, cs / S, g / S, p, raw, st, &src); break;
            }
        }
        if (src != CA_OK) return src;
        if (!staged)
        switch (disp_hw) {
            case 1: launch_raw<1>(fast, fixp, movp, C, h, w, d, cs, g, p, raw, st); break;
            case 2: launch_raw<2>(fast, fixp, movp, C, h, w, d, cs, g, p, raw, st); break;
            case 3: launch_raw<3>(fast, fixp, movp, C, h, w, d, cs, g, p, raw, st); break;
            case 4: launch_raw<4>(fast, fixp, movp, C, h, w, d, cs, g, p, raw, st); break;
            case 5: launch_raw<5>(fast, fixp, movp, C, h, w, d, cs, g, p, raw, st); break;
            case 6: launch_raw<6>(fast, fixp, movp, C, h, w, d, cs, g, p, raw, st); break;
            case 7: launch_raw<7>(fast, fixp, movp, C, h, w, d, cs, g, p, raw, st); break;
            default: launch_raw<8>(fast, fixp, movp, C, h, w, d, cs, g, p, raw, st); break;
        }
        ca_timing_end(tm, st);
        CA_CHECK_LAUNCH("correlate/raw");
        tm = ca_timing_begin(CA_T_BOX, st);
        Box12Args a;
        a.raw = raw; a.ssd = ssd; a.rl = p.rl;
        a.w = w; a.d = d; a.S = S; a.chunk_start = cs;
        a.r0 = p.r0; a.r1 = p.r1; a.b0 = p.b0; a.b1 = p.b1; a.xb = x_begin; a.xe = x_end;
        a.nq = p.nq; a.RP = p.RP; a.Ty = p.Ty;
        const dim3 bgrid((unsigned)p.ntiles, (unsigned)(g * S)), bblock((unsigned)p.nthreads);
        if (fast) {
            BoxFastArgs f;
            f.raw = raw; f.ssd = ssd; f.rl = p.rl;
            f.w = w; f.d = d; f.S = S; f.chunk_start = cs;
            f.r0 = p.r0; f.r1 = p.r1; f.b0 = p.b0; f.b1 = p.b1; f.xb = x_begin; f.xe = x_end;
            f.nq = p.nq; f.RP = p.RP; f.Ty = p.Ty;
            if ((d & 3) == 0) { auto k = k_box12_fast<true>; CA_SMEM_OPTIN(k, 200 * 1024); CA_LAUNCH(k, bgrid, bblock, p.boxf_smem, st, f); }
            else { auto k = k_box12_fast<false>; CA_SMEM_OPTIN(k, 200 * 1024); CA_LAUNCH(k, bgrid, bblock, p.boxf_smem, st, f); }
        } else {
            if ((d & 3) == 0) { auto k = k_box12<true>; CA_SMEM_OPTIN(k, 160 * 1024); CA_LAUNCH(k, bgrid, bblock, p.box_smem, st, a); }
            else { auto k = k_box12<false>; CA_SMEM_OPTIN(k, 160 * 1024); CA_LAUNCH(k, bgrid, bblock, p.box_smem, st, a); }
        }
        ca_timing_end(tm, st);
        CA_CHECK_LAUNCH("correlate/box12");
    }
    if (argmin) {
        unsigned long long* keys = (unsigned long long*)(ws + p.off_keys);
        const int64_t n = (int64_t)(x_end - x_begin) * w * d;
        int rc = ca_launch_cost_argmin(ssd, n, S * S * S, nullptr, nullptr, 0.f, keys, st);
        if (rc != CA_OK) return rc;
        rc = ca_launch_keys_to_i64(keys, n, argmin, st);
        if (rc != CA_OK) return rc;
    }
    return CA_OK;
}

extern "C" int ca_correlate_slab(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, int x_begin,
                                 int x_end, float* ssd, int64_t* argmin, void* workspace, size_t workspace_bytes,
                                 ca_stream_t stream) {
    return correlate_impl(false, fix, mov, C, h, w, d, disp_hw, x_begin, x_end, ssd, argmin, workspace, workspace_bytes, stream);
}

extern "C" int ca_correlate(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, float* ssd,
                            int64_t* argmin, void* workspace, size_t workspace_bytes, ca_stream_t stream) {
    return correlate_impl(false, fix, mov, C, h, w, d, disp_hw, 0, h, ssd, argmin, workspace, workspace_bytes, stream);
}

// ca_correlate_fast for the H planes [x_begin, x_end) only (slab-parallel path)
extern "C" int ca_correlate_slab_fast(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, int x_begin,
                                      int x_end, float* ssd, void* workspace, size_t workspace_bytes, ca_stream_t stream) {
    return correlate_impl(true, fix, mov, C, h, w, d, disp_hw, x_begin, x_end, ssd, nullptr, workspace, workspace_bytes, stream);
}

// The cost volume to a few ulp (relative error <= CA_FAST_SSD_REL_ERR of each entry, see coupled.cu): the input of
// ca_coupled_convex_certified, which still returns the reference's selections bit for bit.
extern "C" int ca_correlate_fast(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, float* ssd,
                                 void* workspace, size_t workspace_bytes, ca_stream_t stream) {
    return correlate_impl(true, fix, mov, C, h, w, d, disp_hw, 0, h, ssd, nullptr, workspace, workspace_bytes, stream);
}
