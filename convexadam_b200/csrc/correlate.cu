// Dense SSD correlation volume over the (2*hw+1)^3 displacement search, with the two 3x3x3 box filters.
//
// Replaces `correlate` (reference src/convexAdam/convex_adam_utils.py:72-89).  The reference materialises an
// im2col of the zero-padded moving features (utils.py:76-77, 9 GB at 192x160x256) and loops over the H-shift;
// here nothing of that exists: labels are produced in chunks of (kH,kW) pairs x all kD,
//     raw  (k_raw_ssd)    : per thread 4 consecutive voxels along D x all S values of kD, channel loop outside
//                           -> every feature load is a 128-bit vector load reused S times from registers
//     box3 (k_box3_exact) : zero-padded 27-tap sum in ascending (dH,dW,dD) order, then one IEEE division by 27
//     box3 again          : writes the chunk's label planes straight into the (L,h,w,d) volume
// with the two chunk-sized intermediates (S planes each, ~51 MB at the north-star shape) recycled so that they stay
// L2-resident while the final volume streams out to HBM once.
//
// EXACT mode: this is a bit-exact restatement of the torch-CPU arithmetic -- (f-m) rounded, squared (x*x)
// rounded, channels accumulated sequentially from c=0, each box3 a sequential fp32 sum of the in-volume taps
// followed by a true division by 27 (avg_pool3d, count_include_pad) -- so that the argmin over labels is the
// reference's argmin.  Every arithmetic op is an explicit _rn intrinsic: nvcc must not contract them to FMAs.
#include "ca_platform.h"

#include <cstdlib>

int ca_launch_cost_argmin(const float* ssd, int64_t n, int L, const float* mesh, const float* soft, float coeff,
                          unsigned long long* keys, cudaStream_t st);
int ca_launch_keys_to_i64(const unsigned long long* keys, int64_t n, int64_t* out, cudaStream_t st);

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

// ---- tile geometry shared by the raw-SSD and box kernels ---------------------------------------------------------
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

// raw[(g*S + kD), x, y, z] = sum_c (f_c(x,y,z) - mpad_c(x+kH, y+kW, z+kD))^2 for the pair (kH,kW) = chunk_start+g
template <int HW, int MINB>
__global__ void __launch_bounds__(128, MINB) k_raw_ssd(const float* __restrict__ fixp, const float* __restrict__ movp,
                                                       int C, int h, int w, int d, int dF, int dP, int chunk_start,
                                                       int zq, int nzb, float* __restrict__ raw) {
    constexpr int S = 2 * HW + 1;
    constexpr int NM = (4 + 2 * HW + 3) & ~3;   // moving values touched by 4 voxels x S shifts, rounded to float4
    const int bz = (int)blockIdx.x % nzb, byi = (int)blockIdx.x / nzb;
    const int q = bz * (int)blockDim.x + (int)threadIdx.x;
    const int y = byi * (int)blockDim.y + (int)threadIdx.y;
    const int x = (int)blockIdx.y;
    if (q >= zq || y >= w) return;
    const int z0 = q * 4;
    const int pi = chunk_start + (int)blockIdx.z;
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
    for (int c = 0; c < C; ++c) {
        const float4 f4 = __ldg(reinterpret_cast<const float4*>(fp));
        const float f[4] = {f4.x, f4.y, f4.z, f4.w};
        float m[NM];
#pragma unroll
        for (int j = 0; j < NM / 4; ++j) {
            const float4 m4 = __ldg(reinterpret_cast<const float4*>(mp) + j);
            m[4 * j + 0] = m4.x; m[4 * j + 1] = m4.y; m[4 * j + 2] = m4.z; m[4 * j + 3] = m4.w;
        }
        fp += fstride;
        mp += mstride;
#pragma unroll
        for (int k = 0; k < S; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float df = __fsub_rn(f[i], m[i + k]);
                acc[k][i] = __fadd_rn(acc[k][i], __fmul_rn(df, df));
            }
    }
    // chunk planes carry a zero halo: (h+2, w+2, dF+8), interior at (+1, +1, +4) -- the box filters then need no
    // bounds checks (the halo is zeroed once per call and never written)
    const int64_t plane = (int64_t)(h + 2) * (w + 2) * (dF + 8);
    float* rp = raw + (int64_t)blockIdx.z * S * plane + ((int64_t)(x + 1) * (w + 2) + (y + 1)) * (dF + 8) + 4 + z0;
#pragma unroll
    for (int k = 0; k < S; ++k) {
        float4 o;
        o.x = (z0 + 0 < d) ? acc[k][0] : 0.f;
        o.y = (z0 + 1 < d) ? acc[k][1] : 0.f;
        o.z = (z0 + 2 < d) ? acc[k][2] : 0.f;
        o.w = (z0 + 3 < d) ? acc[k][3] : 0.f;
        *reinterpret_cast<float4*>(rp + k * plane) = o;
    }
}

// One zero-padded box3 pass over label planes, 4 outputs per thread, ascending (dH,dW,dD) sums, exact /27.
// PADIN : input planes are (h+2, w+2, dF+8) with a zero halo (interior at +1,+1,+4; columns >= d zero): no bounds
//         checks, 9 rows at constant offsets.  !PADIN: plain (h,w,dF) planes, taps outside the volume skipped.
// MODE 0: output plane in the padded layout (pad columns written as zero)
// MODE 1: plane (g*S + kD) goes to label kD*S*S + chunk_start + g of the dense (L,h,w,d) cost volume
// MODE 2: output plane plain (h,w,dF)
template <int MODE, bool PADIN>
__global__ void __launch_bounds__(256) k_box3_exact(const float* __restrict__ in, int h, int w, int d, int dF, int S,
                                                    int chunk_start, int zq, int nzb, float* __restrict__ out) {
    const int bz = (int)blockIdx.x % nzb, byi = (int)blockIdx.x / nzb;
    const int q = bz * (int)blockDim.x + (int)threadIdx.x;
    const int y = byi * (int)blockDim.y + (int)threadIdx.y;
    const int x = (int)blockIdx.y;
    const int l = (int)blockIdx.z;
    if (q >= zq || y >= w) return;
    const int z0 = q * 4;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (PADIN) {
        const int wp = w + 2, dp = dF + 8;
        const float* ip = in + ((int64_t)l * (h + 2) * wp + (int64_t)x * wp + y) * dp + 4 + z0;   // row (x-1, y-1) of the interior
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const float* row = ip + (dx * wp + dy) * dp;
                const float a = row[-1];
                const float4 c = *reinterpret_cast<const float4*>(row);
                const float b = row[4];
                s0 = __fadd_rn(__fadd_rn(__fadd_rn(s0, a), c.x), c.y);
                s1 = __fadd_rn(__fadd_rn(__fadd_rn(s1, c.x), c.y), c.z);
                s2 = __fadd_rn(__fadd_rn(__fadd_rn(s2, c.y), c.z), c.w);
                s3 = __fadd_rn(__fadd_rn(__fadd_rn(s3, c.z), c.w), b);
            }
    } else {
        const float* ip = in + (int64_t)l * h * w * dF + z0;
        const bool has_a = z0 > 0, has_b = z0 + 4 < dF;
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= h) continue;
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= w) continue;
                const float* row = ip + (xx * w + yy) * dF;
                const float4 c = *reinterpret_cast<const float4*>(row);
                const float a = has_a ? row[-1] : 0.f;
                const float b = has_b ? row[4] : 0.f;
                s0 = __fadd_rn(__fadd_rn(__fadd_rn(s0, a), c.x), c.y);
                s1 = __fadd_rn(__fadd_rn(__fadd_rn(s1, c.x), c.y), c.z);
                s2 = __fadd_rn(__fadd_rn(__fadd_rn(s2, c.y), c.z), c.w);
                s3 = __fadd_rn(__fadd_rn(__fadd_rn(s3, c.z), c.w), b);
            }
        }
    }
    s0 = ca_div27(s0); s1 = ca_div27(s1); s2 = ca_div27(s2); s3 = ca_div27(s3);
    if (MODE == 0 || MODE == 2) {
        float4 o;
        o.x = (z0 + 0 < d) ? s0 : 0.f; o.y = (z0 + 1 < d) ? s1 : 0.f; o.z = (z0 + 2 < d) ? s2 : 0.f; o.w = (z0 + 3 < d) ? s3 : 0.f;
        if (MODE == 0)
            *reinterpret_cast<float4*>(out + ((int64_t)l * (h + 2) * (w + 2) + (int64_t)(x + 1) * (w + 2) + (y + 1)) * (dF + 8) + 4 + z0) = o;
        else
            *reinterpret_cast<float4*>(out + ((int64_t)l * h * w + x * w + y) * dF + z0) = o;
    } else {
        const int g = l / S, kD = l - g * S;
        const int64_t label = (int64_t)kD * S * S + chunk_start + g;
        float* op = out + (label * h * w + x * w + y) * d + z0;
        if ((d & 3) == 0) {
            *reinterpret_cast<float4*>(op) = make_float4(s0, s1, s2, s3);
        } else {
            if (z0 + 0 < d) op[0] = s0;
            if (z0 + 1 < d) op[1] = s1;
            if (z0 + 2 < d) op[2] = s2;
            if (z0 + 3 < d) op[3] = s3;
        }
    }
}

// box3 of `nl` planes (h,w,d) with d % 4 == 0 (no pad columns), used by the Adam stage for B^3 (adam.cu)
int ca_launch_box3_planes(const float* in, int nl, int h, int w, int d, float* out, cudaStream_t st) {
    const TileGeom t = make_tile(d, 256);
    auto k = k_box3_exact<2, false>;
    CA_LAUNCH(k, dim3((unsigned)(t.nzb * ((w + t.by - 1) / t.by)), (unsigned)h, (unsigned)nl), dim3(t.bx, t.by), 0, st, in, h, w, d,
              d, 1, 0, t.zq, t.nzb, out);
    CA_CHECK_LAUNCH("box3_planes");
    return CA_OK;
}

namespace {
struct CorrPlan {
    int S, dF, dP, G;
    size_t off_fix, off_mov, off_raw, off_i1, off_keys, total, chunk_bytes;
};

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// G = number of (kH,kW) pairs per chunk; G<=0 -> default (two chunk buffers of ~48 MB: L2 resident on B200)
CorrPlan make_plan(int C, int h, int w, int d, int hw, int G) {
    CorrPlan p;
    p.S = 2 * hw + 1;
    p.dF = ca_roundup4(d);
    p.dP = p.dF + ca_roundup4(4 + 2 * hw) - 4;
    const size_t plane = (size_t)h * w * p.dF * 4;
    if (G <= 0) {
        static int env_g = -1;   // tuning knob: CA_CORR_CHUNK = (kH,kW) pairs per chunk
        if (env_g < 0) { const char* e = getenv("CA_CORR_CHUNK"); env_g = e ? atoi(e) : 0; }
        G = env_g > 0 ? env_g : (int)((48u << 20) / ((size_t)(h + 2) * (w + 2) * (p.dF + 8) * 4 * p.S));
        if (G < 1) G = 1;
    }
    if (G > p.S * p.S) G = p.S * p.S;
    p.G = G;
    size_t o = 0;
    p.off_fix = o; o = align256(o + (size_t)C * plane);
    p.off_mov = o; o = align256(o + (size_t)C * (h + 2 * hw) * (w + 2 * hw) * p.dP * 4);
    const size_t pplane = (size_t)(h + 2) * (w + 2) * (p.dF + 8) * 4;   // chunk planes carry a zero halo
    p.chunk_bytes = (size_t)G * p.S * pplane;
    p.off_raw = o; o = align256(o + p.chunk_bytes);
    p.off_i1 = o; o = align256(o + p.chunk_bytes);
    p.off_keys = o; o = align256(o + (size_t)h * w * d * 8);
    p.total = o;
    return p;
}

static int g_raw_minb = 0;   // tuning knob (CA_RAW_MINB env, read once): min resident blocks per SM of k_raw_ssd

template <int HW>
void launch_raw(const float* fixp, const float* movp, int C, int h, int w, int d, int dF, int dP, int chunk_start, int g,
                float* raw, cudaStream_t st) {
    const TileGeom t = make_tile(dF, 128);
    const dim3 grid((unsigned)(t.nzb * ((w + t.by - 1) / t.by)), (unsigned)h, (unsigned)g), block(t.bx, t.by);
    if (g_raw_minb == 0) {
        const char* e = getenv("CA_RAW_MINB");
        g_raw_minb = e ? atoi(e) : 6;
        if (g_raw_minb != 4 && g_raw_minb != 8) g_raw_minb = 6;
    }
    if (g_raw_minb == 4) { auto k = k_raw_ssd<HW, 4>; CA_LAUNCH(k, grid, block, 0, st, fixp, movp, C, h, w, d, dF, dP, chunk_start, t.zq, t.nzb, raw); }
    else if (g_raw_minb == 8) { auto k = k_raw_ssd<HW, 8>; CA_LAUNCH(k, grid, block, 0, st, fixp, movp, C, h, w, d, dF, dP, chunk_start, t.zq, t.nzb, raw); }
    else { auto k = k_raw_ssd<HW, 6>; CA_LAUNCH(k, grid, block, 0, st, fixp, movp, C, h, w, d, dF, dP, chunk_start, t.zq, t.nzb, raw); }
}
}  // namespace

extern "C" size_t ca_correlate_workspace_bytes(int C, int h, int w, int d, int disp_hw) {
    if (C <= 0 || h <= 0 || w <= 0 || d <= 0 || disp_hw < 1) return 0;
    return make_plan(C, h, w, d, disp_hw, 0).total;
}

extern "C" int ca_correlate(const float* fix, const float* mov, int C, int h, int w, int d, int disp_hw, float* ssd,
                            int64_t* argmin, void* workspace, size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(fix && mov && ssd && workspace, "correlate: null pointer");
    CA_CHECK_ARG(C > 0 && h > 0 && w > 0 && d > 0, "correlate: bad dims");
    if (disp_hw < 1 || disp_hw > 8) {
        ca_set_error("correlate: disp_hw=%d unsupported (1..8)", disp_hw);
        return CA_ERR_UNSUPPORTED;
    }
    CA_CHECK_ARG((int64_t)h * w * ca_roundup4(d) < (1ll << 31) && h < 65536, "correlate: coarse grid too large");
    // largest chunk that fits the caller's workspace (>= 1 pair), capped at the default plan
    CorrPlan p = make_plan(C, h, w, d, disp_hw, 0);
    while (p.total > workspace_bytes && p.G > 1) p = make_plan(C, h, w, d, disp_hw, p.G - 1);
    if (p.total > workspace_bytes) {
        ca_set_error("correlate: workspace too small (%zu < %zu)", workspace_bytes, p.total);
        return CA_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    float* fixp = (float*)(ws + p.off_fix);
    float* movp = (float*)(ws + p.off_mov);
    float* raw = (float*)(ws + p.off_raw);
    float* i1 = (float*)(ws + p.off_i1);
    const int S = p.S, npairs = S * S;
    {
        const int64_t tf = (int64_t)C * h * w * p.dF, tm = (int64_t)C * (h + 2 * disp_hw) * (w + 2 * disp_hw) * p.dP;
        CA_LAUNCH(k_pad_features, dim3((unsigned)ca_min((int)ca_divup(tf, 256), 148 * 16)), dim3(256), 0, st, fix, C, h, w, d, 0, p.dF, fixp);
        CA_LAUNCH(k_pad_features, dim3((unsigned)ca_min((int)ca_divup(tm, 256), 148 * 16)), dim3(256), 0, st, mov, C, h, w, d, disp_hw, p.dP, movp);
        CA_CHECK_LAUNCH("correlate/pad");
        // zero the halos of both chunk buffers once; the kernels only ever write interiors
        if (cudaMemsetAsync(raw, 0, p.chunk_bytes, st) != cudaSuccess || cudaMemsetAsync(i1, 0, p.chunk_bytes, st) != cudaSuccess) {
            ca_set_error("correlate: memset failed");
            return CA_ERR_LAUNCH;
        }
    }
    for (int cs = 0; cs < npairs; cs += p.G) {
        const int g = ca_min(p.G, npairs - cs);
        switch (disp_hw) {
            case 1: launch_raw<1>(fixp, movp, C, h, w, d, p.dF, p.dP, cs, g, raw, st); break;
            case 2: launch_raw<2>(fixp, movp, C, h, w, d, p.dF, p.dP, cs, g, raw, st); break;
            case 3: launch_raw<3>(fixp, movp, C, h, w, d, p.dF, p.dP, cs, g, raw, st); break;
            case 4: launch_raw<4>(fixp, movp, C, h, w, d, p.dF, p.dP, cs, g, raw, st); break;
            case 5: launch_raw<5>(fixp, movp, C, h, w, d, p.dF, p.dP, cs, g, raw, st); break;
            case 6: launch_raw<6>(fixp, movp, C, h, w, d, p.dF, p.dP, cs, g, raw, st); break;
            case 7: launch_raw<7>(fixp, movp, C, h, w, d, p.dF, p.dP, cs, g, raw, st); break;
            default: launch_raw<8>(fixp, movp, C, h, w, d, p.dF, p.dP, cs, g, raw, st); break;
        }
        CA_CHECK_LAUNCH("correlate/raw");
        const int nl = g * S;
        const TileGeom t = make_tile(p.dF, 256);
        const dim3 grid((unsigned)(t.nzb * ((w + t.by - 1) / t.by)), (unsigned)h, (unsigned)nl), block(t.bx, t.by);
        auto k1 = k_box3_exact<0, true>;
        auto k2 = k_box3_exact<1, true>;
        CA_LAUNCH(k1, grid, block, 0, st, raw, h, w, d, p.dF, S, cs, t.zq, t.nzb, i1);
        CA_LAUNCH(k2, grid, block, 0, st, i1, h, w, d, p.dF, S, cs, t.zq, t.nzb, ssd);
        CA_CHECK_LAUNCH("correlate/box3");
    }
    if (argmin) {
        unsigned long long* keys = (unsigned long long*)(ws + p.off_keys);
        const int64_t n = (int64_t)h * w * d;
        int rc = ca_launch_cost_argmin(ssd, n, S * S * S, nullptr, nullptr, 0.f, keys, st);
        if (rc != CA_OK) return rc;
        rc = ca_launch_keys_to_i64(keys, n, argmin, st);
        if (rc != CA_OK) return rc;
    }
    return CA_OK;
}
