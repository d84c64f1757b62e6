/*
 * ca_oracle.c -- CPU restatement of the convexAdam dense-registration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under convexadam_b200/ may import, link or call this file.
 * It is the checker for the CUDA kernels (tests/, __graft_entry__.smoke(), bench.py's
 * cpu_baseline / --impl reference leg) -- never the thing shipped.
 *
 * The arithmetic of the reference lives in third-party torch (unpinned, requirements.txt:5; the
 * container has torch 2.11.0+cu128).  Every function below restates, in plain fp32 C, the torch-CPU
 * semantics at the reference call site it cites (all paths relative to /root/reference/):
 *   orc_avgpool3d          F.avg_pool3d             (src/convexAdam/convex_adam_utils.py:56,85,96,107;
 *                                                    convex_adam_MIND.py:118-119,149-150,166,191)
 *   orc_mindssc            MINDSSC                  (convex_adam_utils.py:24-68)
 *   orc_correlate          correlate                (convex_adam_utils.py:72-89)
 *   orc_coupled_convex     coupled_convex           (convex_adam_utils.py:93-109)
 *   orc_grid_sample3d      F.grid_sample            (convex_adam_utils.py:126-127, convex_adam_MIND.py:174)
 *   orc_inverse_consistency inverse_consistency     (convex_adam_utils.py:114-129)
 *   orc_trilinear          F.interpolate(trilinear) (convex_adam_MIND.py:141,153,182)
 *   orc_adam_*             the Adam loop            (convex_adam_MIND.py:147-181)
 *
 * Parity pinning: tests/golden/make_golden.py runs the UNMODIFIED reference (oracle/ref_shim.py) on
 * seeded inputs and commits its outputs; tests/test_oracle_golden.py checks this file against them
 * (bit-exact for avgpool / correlate / coupled_convex / grid_sample / inverse_consistency,
 * stated tolerances for MIND (libm expf vs SLEEF), trilinear and the Adam gradient).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC  (see oracle/Makefile).
 * -ffp-contract=off matters: every a*b+c below must round twice, as the separate torch ops do.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

ORC_API int orc_version(void) { return 1; }

/* ------------------------------------------------------------------------------------------------
 * avg_pool3d, count_include_pad=True, ceil_mode=False.  torch CPU (avg_pool3d_out_frame): window clipped to
 * the padded extent gives the divisor, window clipped to the real extent is summed sequentially in
 * ascending (t,h,w) order starting from 0, then one true division.
 * in: (nc, iT, iH, iW) contiguous; out: (nc, oT, oH, oW), o = (i + 2*pad - k)/stride + 1.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_avgpool3d(const float* in, int64_t nc, int iT, int iH, int iW, int k, int stride, int pad,
                           float* out) {
    const int oT = (iT + 2 * pad - k) / stride + 1;
    const int oH = (iH + 2 * pad - k) / stride + 1;
    const int oW = (iW + 2 * pad - k) / stride + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t c = 0; c < nc; ++c) {
        for (int t = 0; t < oT; ++t) {
            const float* ip = in + c * (int64_t)iT * iH * iW;
            float* op = out + c * (int64_t)oT * oH * oW;
            for (int i = 0; i < oH; ++i) {
                for (int j = 0; j < oW; ++j) {
                    int ts = t * stride - pad, hs = i * stride - pad, ws = j * stride - pad;
                    int te = imin(ts + k, iT + pad), he = imin(hs + k, iH + pad), we = imin(ws + k, iW + pad);
                    const int pool = (te - ts) * (he - hs) * (we - ws);
                    ts = imax(ts, 0); hs = imax(hs, 0); ws = imax(ws, 0);
                    te = imin(te, iT); he = imin(he, iH); we = imin(we, iW);
                    float sum = 0.f;
                    if (ts < te && hs < he && ws < we) {
                        for (int z = ts; z < te; ++z)
                            for (int y = hs; y < he; ++y)
                                for (int x = ws; x < we; ++x) sum += ip[((int64_t)z * iH + y) * iW + x];
                        sum = sum / (float)pool;
                    }
                    op[((int64_t)t * oH + i) * oW + j] = sum;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * torch CPU `sum` over a strided (non-innermost) dimension of length n (ATen SumKernel.cpp).
 * Columns handled by the 4-vector body accumulate sequentially (cascade levels never trigger for
 * n <= 16); the remaining columns of every loop call go through row_sum(): four interleaved
 * accumulators p[k] = sum_i x[4i+k], leftovers into p[0], then ((p0+p1)+p2)+p3.
 * ---------------------------------------------------------------------------------------------- */
static inline float sum_sequential(const float* x, int n) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += x[i];
    return s;
}
static inline float sum_ilp4(const float* x, int n) {
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    const int q = n / 4;
    for (int i = 0; i < q; ++i)
        for (int k = 0; k < 4; ++k) p[k] += x[4 * i + k];
    for (int i = 4 * q; i < n; ++i) p[0] += x[i];
    for (int k = 1; k < 4; ++k) p[0] += p[k];
    return p[0];
}

/* ------------------------------------------------------------------------------------------------
 * MINDSSC (convex_adam_utils.py:24-68).  img (H,W,D) -> out (12,H,W,D), channel order after the
 * permutation at :66.  Pair table derived exactly as :30-47 does (mask (x>y)&(dist==2), row-major).
 * M (global mean of mind_var) is accumulated in double and rounded to float; torch uses a float
 * cascade sum -- a <=1 ulp difference of the clamp bounds, covered by the stated MIND tolerance.
 * ---------------------------------------------------------------------------------------------- */
static const int SIX[6][3] = {{0, 1, 1}, {1, 1, 0}, {1, 0, 1}, {1, 1, 2}, {2, 1, 1}, {1, 2, 1}};
static const int MIND_PERM[12] = {6, 8, 1, 11, 2, 10, 0, 7, 9, 4, 5, 3};

/* fills a[12][3], b[12][3]: kernel offsets (in units of dilation, range -1..1) of the two shifted
 * reads of OUTPUT channel c (permutation already applied). */
ORC_API void orc_mind_pairs(int* a, int* b) {
    int s1[12][3], s2[12][3], n = 0;
    for (int x = 0; x < 6; ++x)
        for (int y = 0; y < 6; ++y) {
            int dist = 0;
            for (int t = 0; t < 3; ++t) dist += (SIX[x][t] - SIX[y][t]) * (SIX[x][t] - SIX[y][t]);
            if (x > y && dist == 2) {
                for (int t = 0; t < 3; ++t) { s1[n][t] = SIX[x][t] - 1; s2[n][t] = SIX[y][t] - 1; }
                ++n;
            }
        }
    for (int c = 0; c < 12; ++c)
        for (int t = 0; t < 3; ++t) { a[c * 3 + t] = s1[MIND_PERM[c]][t]; b[c * 3 + t] = s2[MIND_PERM[c]][t]; }
}

ORC_API void orc_mindssc(const float* img, int H, int W, int D, int radius, int dil, float* out) {
    int a[36], b[36];
    orc_mind_pairs(a, b);
    const int64_t V = (int64_t)H * W * D;
    const int k = 2 * radius + 1;
    float* dsq = (float*)malloc(sizeof(float) * 12 * V);   /* (shift1 - shift2)^2, utils.py:56 */
    float* var = (float*)malloc(sizeof(float) * V);
#pragma omp parallel for collapse(2) schedule(static)
    for (int c = 0; c < 12; ++c)
        for (int x = 0; x < H; ++x)
            for (int y = 0; y < W; ++y)
                for (int z = 0; z < D; ++z) {
                    const int x1 = clampi(x + dil * a[c * 3 + 0], 0, H - 1), y1 = clampi(y + dil * a[c * 3 + 1], 0, W - 1),
                              z1 = clampi(z + dil * a[c * 3 + 2], 0, D - 1);
                    const int x2 = clampi(x + dil * b[c * 3 + 0], 0, H - 1), y2 = clampi(y + dil * b[c * 3 + 1], 0, W - 1),
                              z2 = clampi(z + dil * b[c * 3 + 2], 0, D - 1);
                    const float df = img[((int64_t)x1 * W + y1) * D + z1] - img[((int64_t)x2 * W + y2) * D + z2];
                    dsq[c * V + ((int64_t)x * W + y) * D + z] = df * df;
                }
    /* replicate-pad(radius) + avg_pool3d(k, stride 1): sequential k^3 sum / k^3 (utils.py:56) */
    double msum = 0.0;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : msum)
    for (int x = 0; x < H; ++x)
        for (int y = 0; y < W; ++y)
            for (int z = 0; z < D; ++z) {
                float s[12];
                for (int c = 0; c < 12; ++c) {
                    float acc = 0.f;
                    for (int dx = -radius; dx <= radius; ++dx)
                        for (int dy = -radius; dy <= radius; ++dy)
                            for (int dz = -radius; dz <= radius; ++dz) {
                                const int xx = clampi(x + dx, 0, H - 1), yy = clampi(y + dy, 0, W - 1),
                                          zz = clampi(z + dz, 0, D - 1);
                                acc += dsq[c * V + ((int64_t)xx * W + yy) * D + zz];
                            }
                    s[c] = acc / (float)(k * k * k);
                }
                /* NB: min/mean run over the un-permuted channels in torch; both are order-independent
                 * up to the summation order of the mean, which we take in un-permuted order. */
                float tmp[12];
                for (int c = 0; c < 12; ++c) tmp[MIND_PERM[c]] = s[c];
                float mn = tmp[0];
                for (int c = 1; c < 12; ++c) mn = tmp[c] < mn ? tmp[c] : mn;
                for (int c = 0; c < 12; ++c) tmp[c] = tmp[c] - mn;
                const float v = sum_sequential(tmp, 12) / 12.f;
                const int64_t o = ((int64_t)x * W + y) * D + z;
                var[o] = v;
                msum += (double)v;
                for (int c = 0; c < 12; ++c) out[c * V + o] = s[c] - mn;
            }
    const float M = (float)(msum / (double)V);
    const float lo = (float)((double)M * 0.001), hi = (float)((double)M * 1000.0);
#pragma omp parallel for schedule(static)
    for (int64_t o = 0; o < V; ++o) {
        float v = var[o];
        v = v < lo ? lo : (v > hi ? hi : v);
        for (int c = 0; c < 12; ++c) out[c * V + o] = expf(-(out[c * V + o] / v));
    }
    free(dsq);
    free(var);
}

/* ------------------------------------------------------------------------------------------------
 * correlate (convex_adam_utils.py:72-89).  fix, mov: (C,h,w,d).  ssd: (L,h,w,d), L = S^3, S = 2hw+1,
 * l = kD*S^2 + kW*S + kH (view/transpose/reshape at :86).  argmin: first minimum (:87).
 *   raw(x)  = sum_c (f_c(x) - mpad_c(x+k))^2, mpad zero outside (F.pad :76); diff and square rounded,
 *             channel sum per torch `sum(0)` (see sum_ilp4 above);
 *   ssd_l   = box3(box3(raw)), box3 = avg_pool3d(3,1,1) zero padded, /27 (:85).
 * `tail`: number of trailing positions of the flattened (h, S*S, w, d) index (per kH iteration of the
 * loop at :83) whose channel sum torch evaluates with row_sum() -- n mod 64 on an AVX-512 host running
 * single threaded, 0 to get the canonical sequential order everywhere (what the CUDA kernel implements).
 * ---------------------------------------------------------------------------------------------- */
static void box3_zero(const float* in, int h, int w, int d, float* out) {
    for (int x = 0; x < h; ++x)
        for (int y = 0; y < w; ++y)
            for (int z = 0; z < d; ++z) {
                float s = 0.f;
                for (int xx = imax(x - 1, 0); xx < imin(x + 2, h); ++xx)
                    for (int yy = imax(y - 1, 0); yy < imin(y + 2, w); ++yy)
                        for (int zz = imax(z - 1, 0); zz < imin(z + 2, d); ++zz) s += in[((int64_t)xx * w + yy) * d + zz];
                out[((int64_t)x * w + y) * d + z] = s / 27.f;
            }
}

ORC_API void orc_correlate(const float* fix, const float* mov, int C, int h, int w, int d, int hw, float* ssd,
                           int64_t* argmin, int64_t tail) {
    const int S = 2 * hw + 1;
    const int64_t n = (int64_t)h * w * d, L = (int64_t)S * S * S;
    const int64_t ninner = (int64_t)h * S * S * w * d;
#pragma omp parallel
    {
        float* raw = (float*)malloc(sizeof(float) * n);
        float* tmp = (float*)malloc(sizeof(float) * n);
        float sq[64];
#pragma omp for schedule(dynamic, 1)
        for (int64_t l = 0; l < L; ++l) {
            const int kH = (int)(l % S), kW = (int)((l / S) % S), kD = (int)(l / ((int64_t)S * S));
            const int s2 = kW * S + kD; /* unfold channel index, utils.py:76-77 */
            for (int x = 0; x < h; ++x)
                for (int y = 0; y < w; ++y)
                    for (int z = 0; z < d; ++z) {
                        const int xm = x + kH - hw, ym = y + kW - hw, zm = z + kD - hw;
                        const int inb = xm >= 0 && xm < h && ym >= 0 && ym < w && zm >= 0 && zm < d;
                        for (int c = 0; c < C; ++c) {
                            const float f = fix[c * n + ((int64_t)x * w + y) * d + z];
                            const float m = inb ? mov[c * n + ((int64_t)xm * w + ym) * d + zm] : 0.f;
                            const float df = f - m;
                            sq[c] = df * df;
                        }
                        const int64_t flat = (((int64_t)x * S * S + s2) * w + y) * d + z;
                        raw[((int64_t)x * w + y) * d + z] = (flat >= ninner - tail) ? sum_ilp4(sq, C) : sum_sequential(sq, C);
                    }
            box3_zero(raw, h, w, d, tmp);
            box3_zero(tmp, h, w, d, ssd + l * n);
        }
        free(raw);
        free(tmp);
    }
    if (argmin) {
#pragma omp parallel for schedule(static)
        for (int64_t v = 0; v < n; ++v) {
            float best = ssd[v];
            int64_t bi = 0;
            for (int64_t l = 1; l < L; ++l) {
                const float s = ssd[l * n + v];
                if (s < best) { best = s; bi = l; }
            }
            argmin[v] = bi;
        }
    }
}

/* One (voxel, label) entry of the cost volume evaluated from its 5x5x5 neighbourhood -- the
 * size-independent checker for full-size volumes (canonical sequential channel order). */
ORC_API float orc_correlate_point(const float* fix, const float* mov, int C, int h, int w, int d, int hw, int64_t l,
                                  int x, int y, int z) {
    const int S = 2 * hw + 1;
    const int64_t n = (int64_t)h * w * d;
    const int kH = (int)(l % S), kW = (int)((l / S) % S), kD = (int)(l / ((int64_t)S * S));
    float raw[5][5][5], i1[3][3][3];
    for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 5; ++b)
            for (int c5 = 0; c5 < 5; ++c5) {
                const int xx = x + a - 2, yy = y + b - 2, zz = z + c5 - 2;
                float r = 0.f;
                if (xx >= 0 && xx < h && yy >= 0 && yy < w && zz >= 0 && zz < d) {
                    const int xm = xx + kH - hw, ym = yy + kW - hw, zm = zz + kD - hw;
                    const int inb = xm >= 0 && xm < h && ym >= 0 && ym < w && zm >= 0 && zm < d;
                    for (int c = 0; c < C; ++c) {
                        const float f = fix[c * n + ((int64_t)xx * w + yy) * d + zz];
                        const float m = inb ? mov[c * n + ((int64_t)xm * w + ym) * d + zm] : 0.f;
                        const float df = f - m;
                        r += df * df;
                    }
                }
                raw[a][b][c5] = r; /* zero outside the volume == skipped tap */
            }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int c3 = 0; c3 < 3; ++c3) {
                const int xx = x + a - 1, yy = y + b - 1, zz = z + c3 - 1;
                float s = 0.f;
                if (xx >= 0 && xx < h && yy >= 0 && yy < w && zz >= 0 && zz < d) {
                    for (int p = 0; p < 3; ++p)
                        for (int q = 0; q < 3; ++q)
                            for (int r = 0; r < 3; ++r) s += raw[a + p][b + q][c3 + r];
                    s = s / 27.f;
                }
                i1[a][b][c3] = s;
            }
    float s = 0.f;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            for (int c3 = 0; c3 < 3; ++c3) s += i1[a][b][c3];
    return s / 27.f;
}

/* ------------------------------------------------------------------------------------------------
 * coupled_convex (convex_adam_utils.py:93-109).  mesh: (3,L) as produced by F.affine_grid at
 * convex_adam_MIND.py:127 (NOT integral -- passed in, never synthesised).  soft: (3,h,w,d) out.
 * sel_out (optional): (7,h,w,d) int64 -- row 0 = the incoming argmin, rows 1..6 = each pass's argmin.
 *   cost_l = ssd_l + coeff * ((m0-s0)^2 + (m1-s1)^2) + (m2-s2)^2)   every op rounded separately;
 *   soft   = box3(mesh[:, sel])  (27-tap zero-padded sum / 27).
 * ---------------------------------------------------------------------------------------------- */
static void soft_from_sel(const float* mesh, int64_t L, const int64_t* sel, int h, int w, int d, float* soft) {
    const int64_t n = (int64_t)h * w * d;
    float* g = (float*)malloc(sizeof(float) * n);
    for (int c = 0; c < 3; ++c) {
        for (int64_t v = 0; v < n; ++v) g[v] = mesh[c * L + sel[v]];
        box3_zero(g, h, w, d, soft + c * n);
    }
    free(g);
}

ORC_API void orc_coupled_convex(const float* ssd, const int64_t* argmin, const float* mesh, int h, int w, int d,
                                int64_t L, float* soft, int64_t* sel_out) {
    static const float coeffs[6] = {0.003f, 0.01f, 0.03f, 0.1f, 0.3f, 1.f};
    const int64_t n = (int64_t)h * w * d;
    int64_t* sel = (int64_t*)malloc(sizeof(int64_t) * n);
    soft_from_sel(mesh, L, argmin, h, w, d, soft);
    if (sel_out) memcpy(sel_out, argmin, sizeof(int64_t) * n);
    for (int j = 0; j < 6; ++j) {
        const float coeff = coeffs[j];
#pragma omp parallel for schedule(static)
        for (int64_t v = 0; v < n; ++v) {
            const float s0 = soft[v], s1 = soft[n + v], s2 = soft[2 * n + v];
            float best = 0.f;
            int64_t bi = 0;
            for (int64_t l = 0; l < L; ++l) {
                const float d0 = mesh[l] - s0, d1 = mesh[L + l] - s1, d2 = mesh[2 * L + l] - s2;
                const float q0 = d0 * d0, q1 = d1 * d1, q2 = d2 * d2;
                float dist = q0 + q1;
                dist = dist + q2;
                const float reg = coeff * dist;
                const float cst = ssd[l * n + v] + reg;
                if (l == 0 || cst < best) { best = cst; bi = l; }
            }
            sel[v] = bi;
        }
        soft_from_sel(mesh, L, sel, h, w, d, soft);
        if (sel_out) memcpy(sel_out + (int64_t)(j + 1) * n, sel, sizeof(int64_t) * n);
    }
    free(sel);
}

/* ------------------------------------------------------------------------------------------------
 * grid_sample, 5-D, mode='bilinear', padding_mode='zeros', align_corners=False (torch CPU
 * grid_sampler_3d_cpu + GridSampler.h:27-36 unnormalise).  in: (C,iD,iH,iW); grid: (oD,oH,oW,3) with
 * grid[...,0]=x (-> iW, last axis), 1=y (-> iH), 2=z (-> iD); out: (C,oD,oH,oW).
 * ---------------------------------------------------------------------------------------------- */
static inline float unnorm(float c, int size) { return ((c + 1.f) * (float)size - 1.f) / 2.f; }

typedef struct {
    int64_t x0, y0, z0;
    float w[8]; /* tnw,tne,tsw,tse,bnw,bne,bsw,bse */
} corner_t;

static inline void corners(float ix, float iy, float iz, corner_t* k) {
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const float x1 = fx + 1.f, y1 = fy + 1.f, z1 = fz + 1.f;
    k->x0 = (int64_t)fx; k->y0 = (int64_t)fy; k->z0 = (int64_t)fz;
    k->w[0] = (x1 - ix) * (y1 - iy) * (z1 - iz);
    k->w[1] = (ix - fx) * (y1 - iy) * (z1 - iz);
    k->w[2] = (x1 - ix) * (iy - fy) * (z1 - iz);
    k->w[3] = (ix - fx) * (iy - fy) * (z1 - iz);
    k->w[4] = (x1 - ix) * (y1 - iy) * (iz - fz);
    k->w[5] = (ix - fx) * (y1 - iy) * (iz - fz);
    k->w[6] = (x1 - ix) * (iy - fy) * (iz - fz);
    k->w[7] = (ix - fx) * (iy - fy) * (iz - fz);
}

ORC_API void orc_grid_sample3d(const float* in, int C, int iD, int iH, int iW, const float* grid, int oD, int oH,
                               int oW, float* out) {
    const int64_t on = (int64_t)oD * oH * oW, in_n = (int64_t)iD * iH * iW;
#pragma omp parallel for schedule(static)
    for (int64_t o = 0; o < on; ++o) {
        const float ix = unnorm(grid[o * 3 + 0], iW), iy = unnorm(grid[o * 3 + 1], iH), iz = unnorm(grid[o * 3 + 2], iD);
        corner_t k;
        corners(ix, iy, iz, &k);
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
            for (int t = 0; t < 8; ++t) {
                const int64_t x = k.x0 + (t & 1), y = k.y0 + ((t >> 1) & 1), z = k.z0 + ((t >> 2) & 1);
                if (x >= 0 && x < iW && y >= 0 && y < iH && z >= 0 && z < iD)
                    acc += in[c * in_n + (z * iH + y) * iW + x] * k.w[t];
            }
            out[c * on + o] = acc;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * inverse_consistency (convex_adam_utils.py:114-129).  f1,f2: (3,h,w,d) normalised, channel 0 = x
 * (last array axis).  gx,gy,gz: identity grid axes of F.affine_grid(eye,(1,1,h,w,d)) (align_corners
 * False), gx has d entries, gy w, gz h.  o1,o2 out.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_inverse_consistency(const float* f1, const float* f2, int h, int w, int d, int iters, const float* gx,
                                     const float* gy, const float* gz, float* o1, float* o2) {
    const int64_t n = (int64_t)h * w * d;
    float* a = (float*)malloc(sizeof(float) * 3 * n);
    float* b = (float*)malloc(sizeof(float) * 3 * n);
    float* grid = (float*)malloc(sizeof(float) * 3 * n);
    float* smp = (float*)malloc(sizeof(float) * 3 * n);
    memcpy(o1, f1, sizeof(float) * 3 * n);
    memcpy(o2, f2, sizeof(float) * 3 * n);
    for (int it = 0; it < iters; ++it) {
        memcpy(a, o1, sizeof(float) * 3 * n);
        memcpy(b, o2, sizeof(float) * 3 * n);
        for (int pass = 0; pass < 2; ++pass) {
            const float* self = pass ? b : a;
            const float* other = pass ? a : b;
            float* dst = pass ? o2 : o1;
            for (int x = 0; x < h; ++x)
                for (int y = 0; y < w; ++y)
                    for (int z = 0; z < d; ++z) {
                        const int64_t o = ((int64_t)x * w + y) * d + z;
                        grid[o * 3 + 0] = gx[z] + self[o];
                        grid[o * 3 + 1] = gy[y] + self[n + o];
                        grid[o * 3 + 2] = gz[x] + self[2 * n + o];
                    }
            orc_grid_sample3d(other, 3, h, w, d, grid, h, w, d, smp);
            for (int64_t i = 0; i < 3 * n; ++i) dst[i] = 0.5f * (self[i] - smp[i]);
        }
    }
    free(a); free(b); free(grid); free(smp);
}

/* ------------------------------------------------------------------------------------------------
 * F.interpolate(mode='trilinear', align_corners=False) (ATen UpSample.h:289-305 source index;
 * UpSampleKernel.cpp cpu_upsample_linear: nested weighted sums, outermost axis first).
 * ---------------------------------------------------------------------------------------------- */
static inline void src_index(int o, int isz, int osz, int* i0, int* i1, float* l0, float* l1) {
    if (isz == osz) { *i0 = o; *i1 = o; *l0 = 1.f; *l1 = 0.f; return; }
    const float scale = (float)isz / (float)osz;
    float s = scale * ((float)o + 0.5f) - 0.5f;
    if (s < 0.f) s = 0.f;
    *i0 = (int)s;
    *i1 = *i0 + ((*i0 < isz - 1) ? 1 : 0);
    *l1 = s - (float)(*i0);
    *l0 = 1.f - *l1;
}

ORC_API void orc_trilinear(const float* in, int C, int ih, int iw, int id, int oh, int ow, int od, float* out) {
    const int64_t in_n = (int64_t)ih * iw * id, on = (int64_t)oh * ow * od;
#pragma omp parallel for collapse(2) schedule(static)
    for (int c = 0; c < C; ++c)
        for (int x = 0; x < oh; ++x) {
            int x0, x1; float a0, a1;
            src_index(x, ih, oh, &x0, &x1, &a0, &a1);
            for (int y = 0; y < ow; ++y) {
                int y0, y1; float b0, b1;
                src_index(y, iw, ow, &y0, &y1, &b0, &b1);
                for (int z = 0; z < od; ++z) {
                    int z0, z1; float c0, c1;
                    src_index(z, id, od, &z0, &z1, &c0, &c1);
                    const float* p = in + c * in_n;
#define AT(X, Y, Z) p[((int64_t)(X) * iw + (Y)) * id + (Z)]
                    const float v =
                        a0 * (b0 * (c0 * AT(x0, y0, z0) + c1 * AT(x0, y0, z1)) + b1 * (c0 * AT(x0, y1, z0) + c1 * AT(x0, y1, z1))) +
                        a1 * (b0 * (c0 * AT(x1, y0, z0) + c1 * AT(x1, y0, z1)) + b1 * (c0 * AT(x1, y1, z0) + c1 * AT(x1, y1, z1)));
#undef AT
                    out[c * on + ((int64_t)x * ow + y) * od + z] = v;
                }
            }
        }
}

/* ------------------------------------------------------------------------------------------------
 * Adam instance optimisation (convex_adam_MIND.py:147-181), one iteration, analytic gradient
 * (SURVEY.md App. A.5).  p,m,v: (3,ha,wa,da) state, updated in place.  fix,mov: (C,ha,wa,da).
 * u_out (3,ha,wa,da) receives B^3(p) of THIS iteration's forward (what the reference returns after
 * its last iteration, :181).  g0x(da), g0y(wa), g0z(ha): identity-grid axes (affine_grid, :160).
 * step: 1-based Adam step count.  g_out (optional) receives d(loss)/dp.
 * ---------------------------------------------------------------------------------------------- */
static void box3_3(const float* in, int h, int w, int d, float* out, float* tmp) {
    const int64_t n = (int64_t)h * w * d;
    for (int c = 0; c < 3; ++c) {
        box3_zero(in + c * n, h, w, d, tmp);
        box3_zero(tmp, h, w, d, out + c * n);
        box3_zero(out + c * n, h, w, d, tmp);
        memcpy(out + c * n, tmp, sizeof(float) * n);
    }
}

ORC_API void orc_adam_forward_u(const float* p, int ha, int wa, int da, float* u) {
    const int64_t n = (int64_t)ha * wa * da;
    float* tmp = (float*)malloc(sizeof(float) * n);
    box3_3(p, ha, wa, da, u, tmp);
    free(tmp);
}

ORC_API void orc_adam_grad(const float* u, const float* fix, const float* mov, int C, int ha, int wa, int da,
                           float lambda_w, const float* g0x, const float* g0y, const float* g0z, float* gu,
                           double* loss_out) {
    const int64_t n = (int64_t)ha * wa * da;
    const float sc[3] = {(float)((ha - 1) / 2.0), (float)((wa - 1) / 2.0), (float)((da - 1) / 2.0)}; /* :170 */
    const double cnt[3] = {3.0 * (ha - 1) * wa * da, 3.0 * ha * (wa - 1) * da, 3.0 * ha * wa * (da - 1)};
    const int dims[3] = {ha, wa, da};
    const int64_t strd[3] = {(int64_t)wa * da, da, 1};
    double loss = 0.0;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : loss)
    for (int x = 0; x < ha; ++x)
        for (int y = 0; y < wa; ++y)
            for (int z = 0; z < da; ++z) {
                const int64_t o = ((int64_t)x * wa + y) * da + z;
                const int idx[3] = {x, y, z};
                /* grid coordinate, exactly the op sequence of :171: grid0 + (u/scale).flip(1) */
                const float gxn = g0x[z] + u[2 * n + o] / sc[2];
                const float gyn = g0y[y] + u[1 * n + o] / sc[1];
                const float gzn = g0z[x] + u[0 * n + o] / sc[0];
                const float ix = unnorm(gxn, da), iy = unnorm(gyn, wa), iz = unnorm(gzn, ha);
                corner_t k;
                corners(ix, iy, iz, &k);
                const float fx = (float)k.x0, fy = (float)k.y0, fz = (float)k.z0;
                const float tx = ix - fx, ty = iy - fy, tz = iz - fz;
                double gix = 0, giy = 0, giz = 0, cost = 0;
                for (int c = 0; c < C; ++c) {
                    float val[8];
                    float acc = 0.f;
                    for (int t = 0; t < 8; ++t) {
                        const int64_t xx = k.x0 + (t & 1), yy = k.y0 + ((t >> 1) & 1), zz = k.z0 + ((t >> 2) & 1);
                        val[t] = (xx >= 0 && xx < da && yy >= 0 && yy < wa && zz >= 0 && zz < ha)
                                     ? mov[c * n + (zz * wa + yy) * da + xx] : 0.f;
                        if (xx >= 0 && xx < da && yy >= 0 && yy < wa && zz >= 0 && zz < ha) acc += val[t] * k.w[t];
                    }
                    const float r = acc - fix[c * n + o];
                    cost += (double)r * r;
                    /* d sample / d ix etc. (trilinear) */
                    const double dx = (1 - ty) * (1 - tz) * ((double)val[1] - val[0]) + ty * (1 - tz) * ((double)val[3] - val[2]) +
                                      (1 - ty) * tz * ((double)val[5] - val[4]) + ty * tz * ((double)val[7] - val[6]);
                    const double dy = (1 - tx) * (1 - tz) * ((double)val[2] - val[0]) + tx * (1 - tz) * ((double)val[3] - val[1]) +
                                      (1 - tx) * tz * ((double)val[6] - val[4]) + tx * tz * ((double)val[7] - val[5]);
                    const double dz = (1 - tx) * (1 - ty) * ((double)val[4] - val[0]) + tx * (1 - ty) * ((double)val[5] - val[1]) +
                                      (1 - tx) * ty * ((double)val[6] - val[2]) + tx * ty * ((double)val[7] - val[3]);
                    gix += r * dx; giy += r * dy; giz += r * dz;
                }
                loss += cost / C * 12.0 / (double)n;
                const double gscale = 2.0 * 12.0 / ((double)C * (double)n);
                /* d pos / d u = (size/2) / scale */
                double g[3];
                g[2] = gscale * gix * ((double)da / 2.0) / sc[2];
                g[1] = gscale * giy * ((double)wa / 2.0) / sc[1];
                g[0] = gscale * giz * ((double)ha / 2.0) / sc[0];
                /* diffusion regulariser (:167-169): lambda * mean((u[i+1]-u[i])^2) per axis over all 3 channels */
                for (int ch = 0; ch < 3; ++ch) {
                    double gr = 0;
                    const double uc = u[ch * n + o];
                    for (int ax = 0; ax < 3; ++ax) {
                        double lap = 0;
                        if (idx[ax] > 0) lap += uc - u[ch * n + o - strd[ax]];
                        if (idx[ax] < dims[ax] - 1) lap += uc - u[ch * n + o + strd[ax]];
                        gr += 2.0 * lambda_w * lap / cnt[ax];
                        if (idx[ax] < dims[ax] - 1) {
                            const double df = (double)u[ch * n + o + strd[ax]] - uc;
                            loss += lambda_w * df * df / cnt[ax];
                        }
                    }
                    gu[ch * n + o] = (float)(g[ch] + gr);
                }
            }
    if (loss_out) *loss_out = loss;
}

ORC_API void orc_adam_update(float* p, float* m, float* v, const float* g, int64_t n3, int step) {
    /* torch/optim/adam.py single-tensor path, lr=1, betas=(0.9,0.999), eps=1e-8, no weight decay */
    const double b1 = 0.9, b2 = 0.999, lr = 1.0, eps = 1e-8;
    const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
    const double step_size = lr / bc1;
    const double bc2_sqrt = sqrt(bc2);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n3; ++i) {
        const float gi = g[i];
        /* exp_avg.lerp_(grad, 1-beta1) ; exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2) */
        m[i] = m[i] + (float)(1.0 - b1) * (gi - m[i]);
        v[i] = v[i] * (float)b2 + (float)(1.0 - b2) * gi * gi;
        const float denom = sqrtf(v[i]) / (float)bc2_sqrt + (float)eps;
        p[i] = p[i] + (float)(-step_size) * (m[i] / denom);
    }
}

ORC_API void orc_adam_step(float* p, float* m, float* v, const float* fix, const float* mov, int C, int ha, int wa,
                           int da, float lambda_w, const float* g0x, const float* g0y, const float* g0z, int step,
                           float* u_out, float* g_out) {
    const int64_t n = (int64_t)ha * wa * da;
    float* gu = (float*)malloc(sizeof(float) * 3 * n);
    float* gp = (float*)malloc(sizeof(float) * 3 * n);
    float* tmp = (float*)malloc(sizeof(float) * n);
    box3_3(p, ha, wa, da, u_out, tmp);
    orc_adam_grad(u_out, fix, mov, C, ha, wa, da, lambda_w, g0x, g0y, g0z, gu, NULL);
    box3_3(gu, ha, wa, da, gp, tmp); /* B is symmetric: dL/dp = B^3 dL/du */
    if (g_out) memcpy(g_out, gp, sizeof(float) * 3 * n);
    orc_adam_update(p, m, v, gp, 3 * n, step);
    free(gu); free(gp); free(tmp);
}
