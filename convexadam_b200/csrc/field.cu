// Coarse-grid field glue kept on the device (no host round trips, unlike the reference's MIND.py:156):
// inverse consistency, trilinear resize, per-channel affine/flip, output packing.
//
//   k_ic_step      inverse_consistency            reference src/convexAdam/convex_adam_utils.py:114-129
//   k_trilinear    F.interpolate(trilinear)       convex_adam_MIND.py:141,153,182   (ATen UpSample.h:289-305)
//   k_field_affine (x/scale).flip(1), *scale*gs   convex_adam_MIND.py:134,139,141
//   k_pack_field   .to(dtype) + np.stack(...,3)   convex_adam_MIND.py:198-202
// All are elementwise / gather kernels over <= a few million voxels: one thread per output voxel, D fastest.
#include "ca_platform.h"

#ifndef CA_EMU
#include <cuda_fp16.h>
#endif

// ---- shared trilinear sampling helpers (torch grid_sampler_3d, bilinear, zeros padding, align_corners=False) ----
struct Corner8 {
    int x0, y0, z0;
    float w[8];   // tnw, tne, tsw, tse, bnw, bne, bsw, bse  (x fastest)
};

static __device__ __forceinline__ float ca_unnormalize(float c, int size) {
    // GridSampler.h:27-36: ((coord + 1) * size - 1) / 2, every op rounded
    return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(c, 1.f), (float)size), 1.f), 0.5f);      // x / 2 == x * 0.5, exactly
}

static __device__ __forceinline__ void ca_corners(float ix, float iy, float iz, Corner8& k) {
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const float x1 = __fadd_rn(fx, 1.f), y1 = __fadd_rn(fy, 1.f), z1 = __fadd_rn(fz, 1.f);
    // clamp before the int conversion: wildly out-of-range coordinates must stay "out of bounds", not overflow
    k.x0 = (int)fminf(fmaxf(fx, -2.f), 1.0e9f);
    k.y0 = (int)fminf(fmaxf(fy, -2.f), 1.0e9f);
    k.z0 = (int)fminf(fmaxf(fz, -2.f), 1.0e9f);
    const float ax = __fsub_rn(x1, ix), bx = __fsub_rn(ix, fx);
    const float ay = __fsub_rn(y1, iy), by = __fsub_rn(iy, fy);
    const float az = __fsub_rn(z1, iz), bz = __fsub_rn(iz, fz);
    k.w[0] = __fmul_rn(__fmul_rn(ax, ay), az);
    k.w[1] = __fmul_rn(__fmul_rn(bx, ay), az);
    k.w[2] = __fmul_rn(__fmul_rn(ax, by), az);
    k.w[3] = __fmul_rn(__fmul_rn(bx, by), az);
    k.w[4] = __fmul_rn(__fmul_rn(ax, ay), bz);
    k.w[5] = __fmul_rn(__fmul_rn(bx, ay), bz);
    k.w[6] = __fmul_rn(__fmul_rn(ax, by), bz);
    k.w[7] = __fmul_rn(__fmul_rn(bx, by), bz);
}

// ---- inverse consistency: one fixed-point iteration for both fields -------------------------------------------
__global__ void __launch_bounds__(256) k_ic_step(const float* __restrict__ a, const float* __restrict__ b, int h, int w,
                                                 int d, const float* __restrict__ gx, const float* __restrict__ gy,
                                                 const float* __restrict__ gz, float* __restrict__ oa,
                                                 float* __restrict__ ob) {
    // the 15 iterations are a chain of short dependent launches: programmatic dependent launch lets the next one's CTAs be
    // scheduled while this one drains; blockIdx.y = which of the two fields this CTA updates (twice the CTAs in flight)
    ca_grid_dep_launch();
    ca_grid_dep_wait();
    const int n = h * w * d;                                   // host: 3 * n < 2^31
    for (int o = (int)(blockIdx.x * blockDim.x + threadIdx.x); o < n; o += (int)(gridDim.x * blockDim.x)) {
        const int r = o / d, z = o - r * d, x = r / w, y = r - x * w;
        const float bgx = gx[z], bgy = gy[y], bgz = gz[x];
        {
            const int pass = (int)blockIdx.y;
            const float* self = pass ? b : a;
            const float* other = pass ? a : b;
            float* dst = pass ? ob : oa;
            const float sx = self[o], sy = self[n + o], sz = self[2 * n + o];
            const float ix = ca_unnormalize(__fadd_rn(bgx, sx), d), iy = ca_unnormalize(__fadd_rn(bgy, sy), w),
                        iz = ca_unnormalize(__fadd_rn(bgz, sz), h);
            Corner8 k;
            ca_corners(ix, iy, iz, k);
            // every corner is loaded (clamped coordinates, all 24 loads in flight); a corner outside the volume gets the
            // weight 0, and acc + v * 0 is acc bit for bit: torch skips it (GridSampler.h within_bounds)
            float acc[3] = {0.f, 0.f, 0.f};
            float val[8][3];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int xx = k.x0 + (t & 1), yy = k.y0 + ((t >> 1) & 1), zz = k.z0 + ((t >> 2) & 1);
                if (!(xx >= 0 && xx < d && yy >= 0 && yy < w && zz >= 0 && zz < h)) k.w[t] = 0.f;
                const int q = (ca_min(ca_max(zz, 0), h - 1) * w + ca_min(ca_max(yy, 0), w - 1)) * d + ca_min(ca_max(xx, 0), d - 1);
#pragma unroll
                for (int c = 0; c < 3; ++c) val[t][c] = __ldg(other + c * n + q);
            }
#pragma unroll
            for (int t = 0; t < 8; ++t)
#pragma unroll
                for (int c = 0; c < 3; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(val[t][c], k.w[t]));
            dst[o] = __fmul_rn(0.5f, __fsub_rn(sx, acc[0]));
            dst[n + o] = __fmul_rn(0.5f, __fsub_rn(sy, acc[1]));
            dst[2 * n + o] = __fmul_rn(0.5f, __fsub_rn(sz, acc[2]));
        }
    }
}

static inline unsigned field_grid(int64_t n) {
    int64_t b = ca_divup(n, 256);
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

extern "C" size_t ca_inverse_consistency_workspace_bytes(int h, int w, int d) { return (size_t)h * w * d * 3 * 4 * 2 + 256; }

extern "C" int ca_inverse_consistency(const float* f1, const float* f2, int h, int w, int d, int iters, const float* gx,
                                      const float* gy, const float* gz, float* o1, float* o2, void* workspace,
                                      size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(f1 && f2 && gx && gy && gz && o1 && o2 && workspace, "inverse_consistency: null pointer");
    CA_CHECK_ARG(h > 0 && w > 0 && d > 0 && iters >= 0, "inverse_consistency: bad dims");
    CA_CHECK_ARG((int64_t)h * w * d * 3 < (1ll << 31), "inverse_consistency: field too large (32-bit voxel indices)");
    if (workspace_bytes < ca_inverse_consistency_workspace_bytes(h, w, d)) {
        ca_set_error("inverse_consistency: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n3 = (int64_t)h * w * d * 3;
    float* wa = (float*)workspace;
    float* wb = wa + n3;
    if (iters == 0) {
        cudaMemcpyAsync(o1, f1, n3 * 4, cudaMemcpyDeviceToDevice, st);
        cudaMemcpyAsync(o2, f2, n3 * 4, cudaMemcpyDeviceToDevice, st);
        return CA_OK;
    }
    const float* sa = f1;
    const float* sb = f2;
    for (int it = 0; it < iters; ++it) {
        const bool to_out = ((iters - it) & 1) == 1;
        float* da = to_out ? o1 : wa;
        float* db = to_out ? o2 : wb;
        CA_LAUNCH_PDL(k_ic_step, dim3(field_grid(n3 / 3), 2u), dim3(256), 0, st, sa, sb, h, w, d, gx, gy, gz, da, db);
        CA_CHECK_LAUNCH("inverse_consistency");
        sa = da;
        sb = db;
    }
    return CA_OK;
}

// ---- per-channel affine + optional channel flip --------------------------------------------------------------
struct Affine3 { float div[3], mul1[3], mul2; int flip; };

__global__ void __launch_bounds__(256) k_field_affine(const float* __restrict__ in, float* __restrict__ out, int64_t n,
                                                      Affine3 p) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < 3 * n; o += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(o / n);
        const int sc = p.flip ? 2 - c : c;
        float v = in[sc * n + (o - c * n)];
        v = ca_div(v, p.div[sc]);
        v = __fmul_rn(v, p.mul1[sc]);
        v = __fmul_rn(v, p.mul2);
        out[o] = v;
    }
}

extern "C" int ca_field_affine(const float* in, float* out, int64_t n_per_channel, int flip, const float div[3],
                               const float mul1[3], float mul2, ca_stream_t stream) {
    CA_CHECK_ARG(in && out && div && mul1 && n_per_channel > 0, "field_affine: bad argument");
    CA_CHECK_ARG(in != out || !flip, "field_affine: in-place flip is not supported");
    Affine3 p;
    for (int i = 0; i < 3; ++i) { p.div[i] = div[i]; p.mul1[i] = mul1[i]; }
    p.mul2 = mul2;
    p.flip = flip;
    CA_LAUNCH(k_field_affine, dim3(field_grid(3 * n_per_channel)), dim3(256), 0, (cudaStream_t)stream, in, out, n_per_channel, p);
    CA_CHECK_LAUNCH("field_affine");
    return CA_OK;
}

// ---- trilinear resize, align_corners=False ----------------------------------------------------------------------
static __device__ __forceinline__ void ca_src_index(int o, int isz, int osz, float scale, int& i0, int& i1, float& l0, float& l1) {
    if (isz == osz) { i0 = o; i1 = o; l0 = 1.f; l1 = 0.f; return; }
    float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)o, 0.5f)), 0.5f);
    s = s < 0.f ? 0.f : s;
    i0 = (int)s;
    i1 = i0 + ((i0 < isz - 1) ? 1 : 0);
    l1 = __fsub_rn(s, (float)i0);
    l0 = __fsub_rn(1.f, l1);
}

__global__ void __launch_bounds__(256) k_trilinear(const float* __restrict__ in, int C, int ih, int iw, int id, int oh,
                                                   int ow, int od, float sh, float sw, float sd, float post_mul,
                                                   float* __restrict__ out) {
    const int64_t on = (int64_t)oh * ow * od, inn = (int64_t)ih * iw * id;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < on; o += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(o % od), y = (int)((o / od) % ow), x = (int)(o / ((int64_t)od * ow));
        int x0, x1, y0, y1, z0, z1;
        float a0, a1, b0, b1, c0, c1;
        ca_src_index(x, ih, oh, sh, x0, x1, a0, a1);
        ca_src_index(y, iw, ow, sw, y0, y1, b0, b1);
        ca_src_index(z, id, od, sd, z0, z1, c0, c1);
        for (int c = 0; c < C; ++c) {
            const float* p = in + c * inn;
#define CA_AT(X, Y, Z) __ldg(p + ((int64_t)(X) * iw + (Y)) * id + (Z))
            const float v00 = __fadd_rn(__fmul_rn(c0, CA_AT(x0, y0, z0)), __fmul_rn(c1, CA_AT(x0, y0, z1)));
            const float v01 = __fadd_rn(__fmul_rn(c0, CA_AT(x0, y1, z0)), __fmul_rn(c1, CA_AT(x0, y1, z1)));
            const float v10 = __fadd_rn(__fmul_rn(c0, CA_AT(x1, y0, z0)), __fmul_rn(c1, CA_AT(x1, y0, z1)));
            const float v11 = __fadd_rn(__fmul_rn(c0, CA_AT(x1, y1, z0)), __fmul_rn(c1, CA_AT(x1, y1, z1)));
#undef CA_AT
            const float v0 = __fadd_rn(__fmul_rn(b0, v00), __fmul_rn(b1, v01));
            const float v1 = __fadd_rn(__fmul_rn(b0, v10), __fmul_rn(b1, v11));
            const float v = __fadd_rn(__fmul_rn(a0, v0), __fmul_rn(a1, v1));
            out[c * on + o] = __fmul_rn(v, post_mul);
        }
    }
}

// The same resize for output rows of a multiple of 4 voxels: a thread owns four consecutive outputs of one row (one 128-bit
// store per channel), the row's H / W source planes and weights are computed once per thread, and nothing is divided in 64 bits.
// Every value is formed by the same rounded operations in the same order as in k_trilinear: the results are bit-identical.
__global__ void __launch_bounds__(256) k_trilinear_q(const float* __restrict__ in, int C, int ih, int iw, int id, int oh,
                                                     int ow, int od, float sh, float sw, float sd, float post_mul,
                                                     float* __restrict__ out) {
    const int nq = od >> 2;
    const int rows_per_cta = (int)blockDim.x / nq;                     // the host picks blockDim.x = a multiple of nq
    const int t = (int)threadIdx.x, rl = t / nq, q = t - rl * nq;
    const int64_t on = (int64_t)oh * ow * od, inn = (int64_t)ih * iw * id;
    for (int r = (int)blockIdx.x * rows_per_cta + rl; r < oh * ow; r += (int)gridDim.x * rows_per_cta) {
        const int x = r / ow, y = r - x * ow;
        int x0, x1, y0, y1;
        float a0, a1, b0, b1;
        ca_src_index(x, ih, oh, sh, x0, x1, a0, a1);
        ca_src_index(y, iw, ow, sw, y0, y1, b0, b1);
        int z0[4], z1[4];
        float c0[4], c1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) ca_src_index(4 * q + j, id, od, sd, z0[j], z1[j], c0[j], c1[j]);
        const int r00 = (x0 * iw + y0) * id, r01 = (x0 * iw + y1) * id, r10 = (x1 * iw + y0) * id, r11 = (x1 * iw + y1) * id;
        for (int c = 0; c < C; ++c) {
            const float* p = in + c * inn;
            float o4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float v00 = __fadd_rn(__fmul_rn(c0[j], __ldg(p + r00 + z0[j])), __fmul_rn(c1[j], __ldg(p + r00 + z1[j])));
                const float v01 = __fadd_rn(__fmul_rn(c0[j], __ldg(p + r01 + z0[j])), __fmul_rn(c1[j], __ldg(p + r01 + z1[j])));
                const float v10 = __fadd_rn(__fmul_rn(c0[j], __ldg(p + r10 + z0[j])), __fmul_rn(c1[j], __ldg(p + r10 + z1[j])));
                const float v11 = __fadd_rn(__fmul_rn(c0[j], __ldg(p + r11 + z0[j])), __fmul_rn(c1[j], __ldg(p + r11 + z1[j])));
                const float v0 = __fadd_rn(__fmul_rn(b0, v00), __fmul_rn(b1, v01));
                const float v1 = __fadd_rn(__fmul_rn(b0, v10), __fmul_rn(b1, v11));
                o4[j] = __fmul_rn(__fadd_rn(__fmul_rn(a0, v0), __fmul_rn(a1, v1)), post_mul);
            }
            *reinterpret_cast<float4*>(out + c * on + (int64_t)r * od + 4 * q) = make_float4(o4[0], o4[1], o4[2], o4[3]);
        }
    }
}

extern "C" int ca_trilinear(const float* in, int C, int ih, int iw, int id, float* out, int oh, int ow, int od,
                            float post_mul, ca_stream_t stream) {
    CA_CHECK_ARG(in && out, "trilinear: null pointer");
    CA_CHECK_ARG(C > 0 && ih > 0 && iw > 0 && id > 0 && oh > 0 && ow > 0 && od > 0, "trilinear: bad dims");
    const float sh = (float)ih / (float)oh, sw = (float)iw / (float)ow, sd = (float)id / (float)od;
    const int nq = od / 4;
    if ((od & 3) == 0 && nq <= 256 && (int64_t)ih * iw * id < (1ll << 31) && (int64_t)oh * ow < (1ll << 31) && (((uintptr_t)out) & 15) == 0 &&
        (((int64_t)oh * ow * od) & 3) == 0) {
        const int threads = (256 / nq) * nq, rows_per_cta = threads / nq;
        int64_t ctas = ca_divup((int64_t)oh * ow, rows_per_cta);
        if (ctas > 148 * 64) ctas = 148 * 64;
        CA_LAUNCH(k_trilinear_q, dim3((unsigned)ctas), dim3(threads), 0, (cudaStream_t)stream, in, C, ih, iw, id, oh, ow, od, sh, sw, sd,
                  post_mul, out);
        CA_CHECK_LAUNCH("trilinear_q");
        return CA_OK;
    }
    CA_LAUNCH(k_trilinear, dim3(field_grid((int64_t)oh * ow * od)), dim3(256), 0, (cudaStream_t)stream, in, C, ih, iw, id, oh,
              ow, od, sh, sw, sd, post_mul, out);
    CA_CHECK_LAUNCH("trilinear");
    return CA_OK;
}

// ---- output packing: (3,H,W,D) f32 -> (H,W,D,3) f64, optionally rounded through half ----------------------------
static __device__ __forceinline__ float ca_round_half(float v) {
#ifdef CA_EMU
    return emu_round_half(v);
#else
    return __half2float(__float2half_rn(v));
#endif
}

__global__ void __launch_bounds__(256) k_pack_field(const float* __restrict__ f, int64_t n, int round_half,
                                                    double* __restrict__ out) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < 3 * n; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = o / 3;
        const int c = (int)(o - 3 * v);
        float x = __ldg(f + c * n + v);
        if (round_half) x = ca_round_half(x);
        out[o] = (double)x;
    }
}

extern "C" int ca_pack_field(const float* field, int H, int W, int D, int round_half, double* out, ca_stream_t stream) {
    CA_CHECK_ARG(field && out && H > 0 && W > 0 && D > 0, "pack_field: bad argument");
    const int64_t n = (int64_t)H * W * D;
    CA_LAUNCH(k_pack_field, dim3(field_grid(3 * n)), dim3(256), 0, (cudaStream_t)stream, field, n, round_half, out);
    CA_CHECK_LAUNCH("pack_field");
    return CA_OK;
}


// ---- apply_convex (reference src/convexAdam/apply_convex.py:13-24) -----------------------------------------------------
// warped = scipy.ndimage.map_coordinates(moving, disp.transpose(3,0,1,2) + identity, order=1): linear interpolation in
// float64 at x + disp(x); a point with any coordinate outside [0, size-1] gets cval = 0 (mode='constant': no
// interpolation beyond the edges), the upper neighbour index is clamped (its weight is then 0).  Output in the dtype of
// `moving`.  disp: (H,W,D,3) float64, the layout convex_adam_pt returns.
template <typename T>
__global__ void __launch_bounds__(256) k_apply_convex(const double* __restrict__ disp, const T* __restrict__ moving, int H, int W,
                                                      int D, T* __restrict__ out) {
    const int64_t n = (int64_t)H * W * D;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(o % D), y = (int)((o / D) % W), x = (int)(o / ((int64_t)D * W));
        const double cx = disp[3 * o] + (double)x, cy = disp[3 * o + 1] + (double)y, cz = disp[3 * o + 2] + (double)z;
        double v = 0.0;
        if (!(cx < 0.0 || cx > (double)(H - 1) || cy < 0.0 || cy > (double)(W - 1) || cz < 0.0 || cz > (double)(D - 1))) {
            const double fx = floor(cx), fy = floor(cy), fz = floor(cz);
            const double tx = cx - fx, ty = cy - fy, tz = cz - fz;
            const int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
            const int x1 = ca_min(x0 + 1, H - 1), y1 = ca_min(y0 + 1, W - 1), z1 = ca_min(z0 + 1, D - 1);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int xi = (c & 4) ? x1 : x0, yi = (c & 2) ? y1 : y0, zi = (c & 1) ? z1 : z0;
                const double w = ((c & 4) ? tx : 1.0 - tx) * ((c & 2) ? ty : 1.0 - ty) * ((c & 1) ? tz : 1.0 - tz);
                v += (double)moving[((int64_t)xi * W + yi) * D + zi] * w;
            }
        }
        out[o] = (T)v;
    }
}

extern "C" int ca_apply_convex(const double* disp, const void* moving, int moving_is_f64, int H, int W, int D, void* warped,
                               ca_stream_t stream) {
    CA_CHECK_ARG(disp && moving && warped && H > 0 && W > 0 && D > 0, "apply_convex: bad argument");
    const int64_t n = (int64_t)H * W * D;
    if (moving_is_f64) {
        auto k = k_apply_convex<double>;
        CA_LAUNCH(k, dim3(field_grid(n)), dim3(256), 0, (cudaStream_t)stream, disp, (const double*)moving, H, W, D, (double*)warped);
    } else {
        auto k = k_apply_convex<float>;
        CA_LAUNCH(k, dim3(field_grid(n)), dim3(256), 0, (cudaStream_t)stream, disp, (const float*)moving, H, W, D, (float*)warped);
    }
    CA_CHECK_LAUNCH("apply_convex");
    return CA_OK;
}


// ---- mean displacement (reference src/convexAdam/convex_adam_translation.py:88-96) -------------------------------------
// translation = np.mean(field[mask], axis=0) (or over the whole volume): three float64 sums and a count, reduced in two
// deterministic levels (per-CTA partials in a fixed order, then one CTA), so the 189 MB field need not leave the device
// when only its mean is wanted.  field: planar (3,n) f32 as the pipeline produces it; mask: optional (n) uint8.
__global__ void __launch_bounds__(256) k_field_sum_partial(const float* __restrict__ field, const unsigned char* __restrict__ mask,
                                                           int64_t n, int round_half, double* __restrict__ partial) {
    __shared__ double red[4][256];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        if (mask && !mask[o]) continue;
        float a = field[o], b = field[n + o], c = field[2 * n + o];
        if (round_half) { a = ca_round_half(a); b = ca_round_half(b); c = ca_round_half(c); }
        s0 += (double)a; s1 += (double)b; s2 += (double)c; cnt += 1.0;
    }
    red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = s2; red[3][threadIdx.x] = cnt;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k)
            for (int j = 0; j < 4; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x < 4) partial[(int64_t)blockIdx.x * 4 + threadIdx.x] = red[threadIdx.x][0];
}
__global__ void __launch_bounds__(256) k_field_sum_final(const double* __restrict__ partial, int nblocks, double* __restrict__ out) {
    __shared__ double red[4][256];
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < nblocks; b += 256)
        for (int j = 0; j < 4; ++j) s[j] += partial[(int64_t)b * 4 + j];
    for (int j = 0; j < 4; ++j) red[j][threadIdx.x] = s[j];
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k)
            for (int j = 0; j < 4; ++j) red[j][threadIdx.x] += red[j][threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double c = red[3][0];
        out[0] = c > 0.0 ? red[0][0] / c : 0.0;
        out[1] = c > 0.0 ? red[1][0] / c : 0.0;
        out[2] = c > 0.0 ? red[2][0] / c : 0.0;
        out[3] = c;
    }
}

extern "C" size_t ca_field_mean_workspace_bytes(void) { return (size_t)1024 * 4 * sizeof(double); }

extern "C" int ca_field_mean(const float* field, const unsigned char* mask, int64_t n, int round_half, double* out4, void* workspace,
                             size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(field && out4 && workspace && n > 0, "field_mean: bad argument");
    if (workspace_bytes < ca_field_mean_workspace_bytes()) {
        ca_set_error("field_mean: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    int64_t nb = ca_divup(n, 256 * 16);
    if (nb > 1024) nb = 1024;
    if (nb < 1) nb = 1;
    double* partial = (double*)workspace;
    CA_LAUNCH(k_field_sum_partial, dim3((unsigned)nb), dim3(256), 0, (cudaStream_t)stream, field, mask, n, round_half, partial);
    CA_LAUNCH(k_field_sum_final, dim3(1), dim3(256), 0, (cudaStream_t)stream, partial, (int)nb, out4);
    CA_CHECK_LAUNCH("field_mean");
    return CA_OK;
}
