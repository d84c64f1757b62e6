// Adam instance optimisation of the dense displacement field.
//
// Replaces the autograd loop of the reference (src/convexAdam/convex_adam_MIND.py:147-181, nnUNet.py:95-133):
// ~70 torch launches + a backward graph per iteration there; here, per iteration,
//     u   = B^3 p                      three exact box3 passes (F.avg_pool3d(3,1,1), MIND.py:166)
//     g_u = dL/du                      ONE kernel: trilinear sample of the moving features at x + u*size/(size-1),
//                                      SSD residual, analytic d/d(position), diffusion regulariser (MIND.py:167-177)
//     g_p = B^3 g_u                    box3 is symmetric (zero padding, /27 everywhere) so backward == forward
//     Adam(p, m, v, g_p)               torch.optim.Adam single-tensor maths (adam.py:528-547), lr=1
// grid_sample's backward is only needed w.r.t. the grid (the features carry no gradient), hence no atomics and a
// deterministic result -- unlike the reference's CUDA path (avg_pool3d_backward uses atomicAdd).
// The scalar loss is never formed: only its gradient enters the update.
#include "box_march.cuh"

#include <cstdlib>

int ca_launch_avgpool3d(const float* in, int64_t nc, int T, int H, int W, int k, int stride, int pad, float* out,
                        cudaStream_t st);

struct AdamGeom {
    int ha, wa, da, C;
    int lo, hl;       // the field arrays hold the H planes [lo, lo+hl) of the grid (whole grid: 0, ha)
    float sc[3];      // (size-1)/2 per axis  (MIND.py:170)
    float gmul[3];    // d(position)/d(u) = (size/2) / sc
    float reg[3];     // 2*lambda / count per axis (MIND.py:167-169)
    float gscale;     // 2 * 12 / (C * N)  (MIND.py:176-177: mean over channels, *12, mean over voxels)
};

// planar (C,n) -> channel-last (n, CG) float4 groups, zero padded to a multiple of 4 channels: the 8-corner gather
// of the moving features then costs 8*CG 128-bit loads per voxel instead of 8*C scalar ones.
__global__ void __launch_bounds__(256) k_planar_to_cl(const float* __restrict__ in, int C, int CG, int64_t n,
                                                      float4* __restrict__ out) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
        for (int cg = 0; cg < CG; ++cg) {
            float4 o;
            o.x = (4 * cg + 0 < C) ? __ldg(in + (int64_t)(4 * cg + 0) * n + v) : 0.f;
            o.y = (4 * cg + 1 < C) ? __ldg(in + (int64_t)(4 * cg + 1) * n + v) : 0.f;
            o.z = (4 * cg + 2 < C) ? __ldg(in + (int64_t)(4 * cg + 2) * n + v) : 0.f;
            o.w = (4 * cg + 3 < C) ? __ldg(in + (int64_t)(4 * cg + 3) * n + v) : 0.f;
            out[v * CG + cg] = o;
        }
}

__global__ void __launch_bounds__(128) k_adam_grad(const float* __restrict__ u, const float4* __restrict__ fixcl,
                                                   const float4* __restrict__ movcl, int CG, AdamGeom g,
                                                   const float* __restrict__ g0x, const float* __restrict__ g0y,
                                                   const float* __restrict__ g0z, float* __restrict__ gu) {
    const int ha = g.ha, wa = g.wa, da = g.da;
    const int n = g.hl * wa * da;                   // field arrays: window-local indexing
    const int z = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int)blockIdx.y, xl = (int)blockIdx.z, x = xl + g.lo;
    if (z >= da) return;
    const int o = (xl * wa + y) * da + z;
    const int of = (x * wa + y) * da + z;           // feature arrays: whole-grid indexing
    const float u0 = u[o], u1 = u[n + o], u2 = u[2 * n + o];
    // diffusion regulariser first (its 18 neighbour loads are independent of everything else and go out together
    // with u): d/du_c(x) of lambda * mean_k (u[x+e_k]-u[x])^2, one-sided at the faces
    float greg[3];
    {
        const int idx[3] = {x, y, z};
        const int dims[3] = {ha, wa, da};
        const int strd[3] = {wa * da, da, 1};
        const float uc[3] = {u0, u1, u2};
        float lo[3][3], hi[3][3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                // a window edge that is not a face of the grid has no neighbour plane: that row of the window is margin
                // the caller discards (slab.py), it only must not be read out of bounds
                const bool has_lo = ax == 0 ? xl > 0 : idx[ax] > 0, has_hi = ax == 0 ? xl < g.hl - 1 : idx[ax] < dims[ax] - 1;
                lo[ch][ax] = has_lo ? u[ch * n + o - strd[ax]] : uc[ch];
                hi[ch][ax] = has_hi ? u[ch * n + o + strd[ax]] : uc[ch];
            }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float gr = 0.f;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                float lap = 0.f;
                if (idx[ax] > 0) lap += uc[ch] - lo[ch][ax];
                if (idx[ax] < dims[ax] - 1) lap += uc[ch] - hi[ch][ax];
                gr += g.reg[ax] * lap;
            }
            greg[ch] = gr;
        }
    }
    // grid_disp = grid0 + (u/scale).flip(1)   (MIND.py:171), then GridSampler.h:27-36, every op rounded as torch does
    const float cx = __fadd_rn(g0x[z], ca_div(u2, g.sc[2]));
    const float cy = __fadd_rn(g0y[y], ca_div(u1, g.sc[1]));
    const float cz = __fadd_rn(g0z[x], ca_div(u0, g.sc[0]));
    const float ix = ca_div(__fsub_rn(__fmul_rn(__fadd_rn(cx, 1.f), (float)da), 1.f), 2.f);
    const float iy = ca_div(__fsub_rn(__fmul_rn(__fadd_rn(cy, 1.f), (float)wa), 1.f), 2.f);
    const float iz = ca_div(__fsub_rn(__fmul_rn(__fadd_rn(cz, 1.f), (float)ha), 1.f), 2.f);
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const float tx = __fsub_rn(ix, fx), ty = __fsub_rn(iy, fy), tz = __fsub_rn(iz, fz);
    const float sx = __fsub_rn(__fadd_rn(fx, 1.f), ix), sy = __fsub_rn(__fadd_rn(fy, 1.f), iy), sz = __fsub_rn(__fadd_rn(fz, 1.f), iz);
    const int x0 = (int)fminf(fmaxf(fx, -2.f), 1.0e9f), y0 = (int)fminf(fmaxf(fy, -2.f), 1.0e9f), z0 = (int)fminf(fmaxf(fz, -2.f), 1.0e9f);
    float wgt[8];
    wgt[0] = __fmul_rn(__fmul_rn(sx, sy), sz); wgt[1] = __fmul_rn(__fmul_rn(tx, sy), sz);
    wgt[2] = __fmul_rn(__fmul_rn(sx, ty), sz); wgt[3] = __fmul_rn(__fmul_rn(tx, ty), sz);
    wgt[4] = __fmul_rn(__fmul_rn(sx, sy), tz); wgt[5] = __fmul_rn(__fmul_rn(tx, sy), tz);
    wgt[6] = __fmul_rn(__fmul_rn(sx, ty), tz); wgt[7] = __fmul_rn(__fmul_rn(tx, ty), tz);
    int off[8];
    bool inb[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int xx = x0 + (t & 1), yy = y0 + ((t >> 1) & 1), zz = z0 + ((t >> 2) & 1);
        inb[t] = xx >= 0 && xx < da && yy >= 0 && yy < wa && zz >= 0 && zz < ha;
        off[t] = inb[t] ? ((zz * wa + yy) * da + xx) * CG : 0;
    }
    // d(w_t)/d(ix) = +-(Y_t Z_t) etc.: sum_t val_t dw_t = sum over corner pairs of (val_hi - val_lo) * (product of the other two)
    const float yz[4] = {sy * sz, ty * sz, sy * tz, ty * tz};
    const float xz[4] = {sx * sz, tx * sz, sx * tz, tx * tz};
    const float xy[4] = {sx * sy, tx * sy, sx * ty, tx * ty};
    float gix = 0.f, giy = 0.f, giz = 0.f;
    // the 8-corner gather of channel group cg+1 is in flight while group cg is reduced
    float4 val[8], nval[8], f, nf;
#pragma unroll
    for (int t = 0; t < 8; ++t) nval[t] = inb[t] ? __ldg(movcl + off[t]) : make_float4(0.f, 0.f, 0.f, 0.f);
    nf = __ldg(fixcl + of * CG);
    for (int cg = 0; cg < CG; ++cg) {
#pragma unroll
        for (int t = 0; t < 8; ++t) val[t] = nval[t];
        f = nf;
        if (cg + 1 < CG) {
#pragma unroll
            for (int t = 0; t < 8; ++t) nval[t] = inb[t] ? __ldg(movcl + off[t] + cg + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
            nf = __ldg(fixcl + of * CG + cg + 1);
        }
        float4 smp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (inb[t]) {   // sample = sum over the in-bounds corners in torch's order (tnw, tne, tsw, tse, bnw, ...)
                smp.x = __fadd_rn(smp.x, __fmul_rn(val[t].x, wgt[t]));
                smp.y = __fadd_rn(smp.y, __fmul_rn(val[t].y, wgt[t]));
                smp.z = __fadd_rn(smp.z, __fmul_rn(val[t].z, wgt[t]));
                smp.w = __fadd_rn(smp.w, __fmul_rn(val[t].w, wgt[t]));
            }
        }
        const float rx = smp.x - f.x, ry = smp.y - f.y, rz = smp.z - f.z, rw = smp.w - f.w;
        float dot[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) dot[t] = rx * val[t].x + ry * val[t].y + rz * val[t].z + rw * val[t].w;
        gix += yz[0] * (dot[1] - dot[0]) + yz[1] * (dot[3] - dot[2]) + yz[2] * (dot[5] - dot[4]) + yz[3] * (dot[7] - dot[6]);
        giy += xz[0] * (dot[2] - dot[0]) + xz[1] * (dot[3] - dot[1]) + xz[2] * (dot[6] - dot[4]) + xz[3] * (dot[7] - dot[5]);
        giz += xy[0] * (dot[4] - dot[0]) + xy[1] * (dot[5] - dot[1]) + xy[2] * (dot[6] - dot[2]) + xy[3] * (dot[7] - dot[3]);
    }
    gu[2 * n + o] = g.gscale * gix * g.gmul[2] + greg[2];
    gu[n + o] = g.gscale * giy * g.gmul[1] + greg[1];
    gu[o] = g.gscale * giz * g.gmul[0] + greg[0];
}

struct AdamScalars { float w_lerp, beta2, one_m_beta2, bc2_sqrt, eps, neg_step_size; };

// torch.optim.Adam single-tensor step (adam.py:528-547), every op rounded as torch's kernels round it
static __device__ __forceinline__ void adam_update1(float& p, float& m, float& v, float g, const AdamScalars& s) {
    m = __fadd_rn(m, __fmul_rn(s.w_lerp, __fsub_rn(g, m)));                                   // exp_avg.lerp_(grad, 1-beta1)
    v = __fadd_rn(__fmul_rn(v, s.beta2), __fmul_rn(__fmul_rn(s.one_m_beta2, g), g));          // mul_(beta2).addcmul_
    const float denom = __fadd_rn(ca_div(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
    p = __fadd_rn(p, __fmul_rn(s.neg_step_size, ca_div(m, denom)));                           // addcdiv_(value=-step_size)
}

__global__ void __launch_bounds__(256) k_adam_update(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                     const float* __restrict__ g, int64_t n3, AdamScalars s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
        float pi = p[i], mi = m[i], vi = v[i];
        adam_update1(pi, mi, vi, g[i], s);
        m[i] = mi; v[i] = vi; p[i] = pi;
    }
}

// ---- one B pass (F.avg_pool3d(x, 3, stride=1, padding=1), MIND.py:166) as a march along H ---------------------------
// in/out: (3, h, w, d) with d % 4 == 0.  A thread owns two rows x four voxels and a segment of Lx planes; it reads the
// planes xs-1 .. xe once each (planes/rows outside the volume are an all-zero row) and keeps three running sums, see
// box_march.cuh.  No shared memory, no block synchronisation.  UPDATE: the finished value is dL/dp, and instead of
// being stored it drives the Adam step of (p, m, v) in place -- the last pass of B^3 g_u and the update in one kernel.
struct FieldBoxArgs {
    const float* in;
    float* out;
    const float* zero_row;
    int h, w, d, nq, Lx, rpb;     // rpb = row pairs per block
    float *p, *m, *v;
    AdamScalars s;
};

// per-thread state of the march: cp[j] = the thread's row j of the CURRENT plane (rows that do not exist point at the
// all-zero row and have stride 0), eoff = element offset of the thread's first output row in plane X-1
struct FieldBoxThread {
    const float* cp[4];
    int64_t stride[4];
    int64_t eoff, plane;
    BoxLane bl;
    bool ook[2];
};

template <int R, bool UPDATE>
static __device__ __forceinline__ void field_box_step(int X, int xs, int xe, const FieldBoxArgs& a, FieldBoxThread& t,
                                                      float (&acc)[3][2][4], float4 (&cur)[4]) {
    constexpr int Cs = (R + 1) % 3;
    if (X > xe) return;
    // next plane first: its latency hides behind this plane's additions (planes outside the volume are zero)
    float4 nxt[4];
    if (X + 1 < a.h && X + 1 <= xe) {
#pragma unroll
        for (int j = 0; j < 4; ++j) nxt[j] = __ldg(reinterpret_cast<const float4*>(t.cp[j] + t.stride[j]));
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) nxt[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (X >= 0 && X < a.h) {
        float hl[4], hr[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) box_halo(cur[j], t.cp[j], t.bl, hl[j], hr[j]);
        box_add_plane<R>(acc, cur, hl, hr);
    } else {
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[R][o][k] = 0.f;      // a zero plane: the fresh sum starts at 0, the others keep their value
    }
    if (X - 1 >= xs) {
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            if (!t.ook[o]) continue;
            const int64_t off = t.eoff + (int64_t)o * a.d;
            float4 r;
            r.x = ca_div27(acc[Cs][o][0]); r.y = ca_div27(acc[Cs][o][1]); r.z = ca_div27(acc[Cs][o][2]); r.w = ca_div27(acc[Cs][o][3]);
            if (UPDATE) {
                float4 pp = *reinterpret_cast<const float4*>(a.p + off), mm = *reinterpret_cast<const float4*>(a.m + off),
                       vv = *reinterpret_cast<const float4*>(a.v + off);
                adam_update1(pp.x, mm.x, vv.x, r.x, a.s); adam_update1(pp.y, mm.y, vv.y, r.y, a.s);
                adam_update1(pp.z, mm.z, vv.z, r.z, a.s); adam_update1(pp.w, mm.w, vv.w, r.w, a.s);
                *reinterpret_cast<float4*>(a.p + off) = pp;
                *reinterpret_cast<float4*>(a.m + off) = mm;
                *reinterpret_cast<float4*>(a.v + off) = vv;
                if (a.out) *reinterpret_cast<float4*>(a.out + off) = r;
            } else {
                *reinterpret_cast<float4*>(a.out + off) = r;
            }
        }
    }
    t.eoff += t.plane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        t.cp[j] += t.stride[j];
        cur[j] = nxt[j];
    }
}

// grid = (row-pair blocks, H segments, 3 channels); block = nq * rpb threads rounded up to whole warps, z-fastest
template <bool UPDATE>
__global__ void __launch_bounds__(256, 1) k_field_box(FieldBoxArgs a) {
    const int tid = (int)threadIdx.x;
    const int q = tid % a.nq, rpl = tid / a.nq;
    const int lane = tid & 31;
    const int rp = (int)blockIdx.x * a.rpb + rpl;
    const int y0 = 2 * rp;
    const bool active = rpl < a.rpb && y0 < a.w;
    const int xs = (int)blockIdx.y * a.Lx, xe = ca_min(a.h, xs + a.Lx);
    FieldBoxThread t;
    t.plane = (int64_t)a.w * a.d;
    const int64_t chan = (int64_t)blockIdx.z * a.h * t.plane;
    t.bl.first = q == 0;
    t.bl.last = q == a.nq - 1;
    t.bl.fb_l = !t.bl.first && lane == 0;
    t.bl.fb_r = !t.bl.last && lane == 31;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = y0 - 1 + j;
        const bool ok = active && y >= 0 && y < a.w;
        t.stride[j] = ok ? t.plane : 0;
        t.cp[j] = ok ? a.in + chan + (int64_t)(xs - 1) * t.plane + (int64_t)y * a.d + 4 * q : a.zero_row + 4;
    }
    t.eoff = chan + (int64_t)(xs - 2) * t.plane + (int64_t)y0 * a.d + 4 * q;     // plane X-1 at the first step X = xs-1
    t.ook[0] = active;
    t.ook[1] = active && y0 + 1 < a.w;
    float acc[3][2][4];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][o][k] = 0.f;
    float4 cur[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        cur[j] = xs >= 1 ? __ldg(reinterpret_cast<const float4*>(t.cp[j])) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int X = xs - 1; X <= xe; X += 3) {
        field_box_step<0, UPDATE>(X, xs, xe, a, t, acc, cur);
        field_box_step<1, UPDATE>(X + 1, xs, xe, a, t, acc, cur);
        field_box_step<2, UPDATE>(X + 2, xs, xe, a, t, acc, cur);
    }
}

static inline unsigned adam_grid(int64_t n, int block) {
    int64_t b = ca_divup(n, block);
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

static int g_field_lx = 0;   // tuning knob (CA_FIELD_LX env, read once): planes per H segment of k_field_box

// one B pass on a 3-channel field; with `upd` the result is consumed by the Adam step (and stored only if out != null)
static int field_box(const float* in, float* out, int ha, int wa, int da, const float* zero_row, const FieldBoxArgs* upd,
                     cudaStream_t st) {
    FieldBoxArgs a;
    if (upd) a = *upd;
    a.in = in; a.out = out; a.zero_row = zero_row;
    a.h = ha; a.w = wa; a.d = da; a.nq = da / 4;
    if (g_field_lx == 0) {
        const char* e = getenv("CA_FIELD_LX");
        g_field_lx = e ? atoi(e) : 8;
        if (g_field_lx < 1) g_field_lx = 8;
    }
    a.Lx = g_field_lx;
    a.rpb = a.nq >= 128 ? 1 : 128 / a.nq;
    const int rps = (wa + 1) / 2;
    if (a.rpb > rps) a.rpb = rps;
    const unsigned nthreads = (unsigned)((a.nq * a.rpb + 31) & ~31);
    const dim3 grid((unsigned)((rps + a.rpb - 1) / a.rpb), (unsigned)((ha + a.Lx - 1) / a.Lx), 3u);
    if (upd) { auto k = k_field_box<true>; CA_LAUNCH(k, grid, dim3(nthreads), 0, st, a); }
    else { auto k = k_field_box<false>; CA_LAUNCH(k, grid, dim3(nthreads), 0, st, a); }
    CA_CHECK_LAUNCH("field_box");
    return CA_OK;
}

extern "C" size_t ca_adam_window_workspace_bytes(int C, int ha, int wa, int da, int lo, int hi) {
    if (C <= 0 || ha <= 0 || wa <= 0 || da <= 0 || lo < 0 || hi > ha || lo >= hi) return 0;
    const size_t n = (size_t)ha * wa * da, nw = (size_t)(hi - lo) * wa * da;
    return nw * 3 * 4 * 3 + 2 * n * (size_t)((C + 3) & ~3) * 4 + 1024 + ((size_t)da + 64) * 4;
}
extern "C" size_t ca_adam_workspace_bytes(int C, int ha, int wa, int da) { return ca_adam_window_workspace_bytes(C, ha, wa, da, 0, ha); }

extern "C" int ca_adam_run_window(const float* fix, const float* mov, int C, int ha, int wa, int da, int lo, int hi,
                                  float lambda_weight, float* p, float* m, float* v, int first_step, int niter, const float* g0x,
                                  const float* g0y, const float* g0z, float* u_out, float* g_out, void* workspace,
                                  size_t workspace_bytes, int features_ready, ca_stream_t stream) {
    CA_CHECK_ARG(fix && mov && p && m && v && g0x && g0y && g0z && u_out && workspace, "adam_run: null pointer");
    CA_CHECK_ARG(C > 0 && ha > 1 && wa > 1 && da > 1, "adam_run: bad dims (each Adam-grid axis must be >= 2)");
    CA_CHECK_ARG(lo >= 0 && hi <= ha && lo < hi, "adam_run: bad window [%d, %d) of %d planes", lo, hi, ha);
    CA_CHECK_ARG(first_step >= 1 && niter >= 0, "adam_run: bad step range");
    CA_CHECK_ARG((int64_t)ha * wa * da * ((C + 3) / 4) < (1ll << 31) && wa < 65536 && ha < 65536, "adam_run: grid too large");
    if (workspace_bytes < ca_adam_window_workspace_bytes(C, ha, wa, da, lo, hi)) {
        ca_set_error("adam_run: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int hl = hi - lo;
    const int64_t n = (int64_t)ha * wa * da, nw = (int64_t)hl * wa * da;
    float* t0 = (float*)workspace;
    float* t1 = t0 + 3 * nw;
    float* gu = t1 + 3 * nw;
    const int CG = (C + 3) / 4;
    float4* fixcl = (float4*)(((uintptr_t)(gu + 3 * nw) + 255) & ~(uintptr_t)255);
    float4* movcl = fixcl + n * CG;
    float* zero_row = (float*)(((uintptr_t)(movcl + n * CG) + 255) & ~(uintptr_t)255);   // da + 16 zeros for k_field_box
    // marching box passes need float4 rows; any other shape takes the generic avg_pool3d kernel
    const bool fast = (da & 3) == 0 && da / 4 <= 256 && (((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)u_out | (uintptr_t)g_out) & 15) == 0;
    if (niter > 0 && fast && cudaMemsetAsync(zero_row, 0, ((size_t)da + 16) * 4, st) != cudaSuccess) {
        ca_set_error("adam_run: memset failed");
        return CA_ERR_LAUNCH;
    }
    auto box = [&](const float* in, float* out) -> int {
        if (fast) return field_box(in, out, hl, wa, da, zero_row, nullptr, st);
        return ca_launch_avgpool3d(in, 3, hl, wa, da, 3, 1, 1, out, st);
    };
    if (niter > 0 && !features_ready) {
        const unsigned gb = adam_grid(n, 256);
        CA_LAUNCH(k_planar_to_cl, dim3(gb), dim3(256), 0, st, fix, C, CG, n, fixcl);
        CA_LAUNCH(k_planar_to_cl, dim3(gb), dim3(256), 0, st, mov, C, CG, n, movcl);
        CA_CHECK_LAUNCH("adam/planar_to_cl");
    }
    AdamGeom g;
    g.ha = ha; g.wa = wa; g.da = da; g.C = C; g.lo = lo; g.hl = hl;
    const int dims[3] = {ha, wa, da};
    const double cnt[3] = {3.0 * (ha - 1) * wa * da, 3.0 * ha * (wa - 1) * da, 3.0 * ha * wa * (da - 1)};
    for (int a = 0; a < 3; ++a) {
        g.sc[a] = (float)((dims[a] - 1) / 2.0);
        g.gmul[a] = (float)((dims[a] / 2.0) / (double)g.sc[a]);
        g.reg[a] = (float)(2.0 * (double)lambda_weight / cnt[a]);
    }
    g.gscale = (float)(2.0 * 12.0 / ((double)C * (double)n));
    const double b1 = 0.9, b2 = 0.999, lr = 1.0, eps = 1e-8;
    for (int it = 0; it < niter; ++it) {
        const int step = first_step + it;
        int rc;
        if ((rc = box(p, t0)) != CA_OK) return rc;
        if ((rc = box(t0, t1)) != CA_OK) return rc;
        if ((rc = box(t1, u_out)) != CA_OK) return rc;
        const int bx = da >= 128 ? 128 : (da >= 64 ? 64 : 32);
        CA_LAUNCH(k_adam_grad, dim3((unsigned)((da + bx - 1) / bx), (unsigned)wa, (unsigned)hl), dim3(bx), 0, st, u_out, fixcl, movcl, CG,
                  g, g0x, g0y, g0z, gu);
        CA_CHECK_LAUNCH("adam_grad");
        if ((rc = box(gu, t0)) != CA_OK) return rc;
        if ((rc = box(t0, t1)) != CA_OK) return rc;
        FieldBoxArgs upd;
        const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
        upd.p = p; upd.m = m; upd.v = v;
        upd.s.w_lerp = (float)(1.0 - b1);
        upd.s.beta2 = (float)b2;
        upd.s.one_m_beta2 = (float)(1.0 - b2);
        upd.s.bc2_sqrt = (float)sqrt(bc2);
        upd.s.eps = (float)eps;
        upd.s.neg_step_size = (float)(-(lr / bc1));
        if (fast) {
            // third pass of B^3 g_u fused with the Adam step; dL/dp is stored only when the caller asks for it
            if ((rc = field_box(t1, g_out, hl, wa, da, zero_row, &upd, st)) != CA_OK) return rc;
        } else {
            float* gp = g_out ? g_out : t0;
            if ((rc = box(t1, gp)) != CA_OK) return rc;
            CA_LAUNCH(k_adam_update, dim3(adam_grid(3 * nw, 256)), dim3(256), 0, st, p, m, v, gp, 3 * nw, upd.s);
            CA_CHECK_LAUNCH("adam_update");
        }
    }
    return CA_OK;
}

extern "C" int ca_adam_run(const float* fix, const float* mov, int C, int ha, int wa, int da, float lambda_weight, float* p,
                           float* m, float* v, int first_step, int niter, const float* g0x, const float* g0y,
                           const float* g0z, float* u_out, float* g_out, void* workspace, size_t workspace_bytes,
                           ca_stream_t stream) {
    return ca_adam_run_window(fix, mov, C, ha, wa, da, 0, ha, lambda_weight, p, m, v, first_step, niter, g0x, g0y, g0z, u_out,
                              g_out, workspace, workspace_bytes, 0, stream);
}
