// Adam instance optimisation of the dense displacement field.
//
// Replaces the autograd loop of the reference (src/convexAdam/convex_adam_MIND.py:147-181, nnUNet.py:95-133):
// ~70 torch launches + a backward graph per iteration there; here, per iteration,
//     u   = B^3 p                      three box3 passes (F.avg_pool3d(3,1,1), MIND.py:166) fused into two separable kernels
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

// L2 eviction priorities for the gathers (tuning switch, off): the fixed features are read once per voxel and iteration and
// could be loaded evict-first, the moving features and fields evict-last.  Measured at config 2 (100 iterations): 8.90 ms
// without hints, 9.17 ms with evict-first fixed features, 9.20 ms with both -- features + fields (165 MB) exceed the 126 MB L2
// either way and the hinted loads cost issue slots.  CA_GRAD_HINTS=0/1/2 = none / fixed only / fixed + moving.
#ifndef CA_GRAD_HINTS
#define CA_GRAD_HINTS 0
#endif
#if CA_GRAD_HINTS >= 1
#define CA_GRAD_LD_FIX(p) ca_ld_hint4((p), pol_first)
#else
#define CA_GRAD_LD_FIX(p) __ldg(p)
#endif
#if CA_GRAD_HINTS >= 2
#define CA_GRAD_LD_MOV(p) ca_ld_hint4((p), pol_last)
#else
#define CA_GRAD_LD_MOV(p) __ldg(p)
#endif
struct AdamGeom {
    int ha, wa, da, C;
    int lo, hl;       // the field arrays hold the H planes [lo, lo+hl) of the grid (whole grid: 0, ha)
    float sc[3];      // (size-1)/2 per axis  (MIND.py:170)
    float gmul[3];    // d(position)/d(u) = (size/2) / sc
    float reg[3];     // 2*lambda / count per axis (MIND.py:167-169)
    float gscale;     // 2 * 12 / (C * N)  (MIND.py:176-177: mean over channels, *12, mean over voxels)
    int vs, gs;       // feature groups: float4 index = voxel * vs + group * gs  ((n, CG): vs = CG, gs = 1; (CG, n): vs = 1, gs = n)
    int flip;         // 1: the CTAs walk the H planes downwards (alternate iterations: the planes touched last are still in L2)
};

// planar (C,n) -> channel-last (n, CG) float4 groups, zero padded to a multiple of 4 channels: the 8-corner gather
// of the moving features then costs 8*CG 128-bit loads per voxel instead of 8*C scalar ones.
__global__ void __launch_bounds__(256) k_planar_to_cl(const float* __restrict__ in, int C, int CG, int64_t n, int64_t vs, int64_t gs,
                                                      float4* __restrict__ out) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
        for (int cg = 0; cg < CG; ++cg) {
            float4 o;
            o.x = (4 * cg + 0 < C) ? __ldg(in + (int64_t)(4 * cg + 0) * n + v) : 0.f;
            o.y = (4 * cg + 1 < C) ? __ldg(in + (int64_t)(4 * cg + 1) * n + v) : 0.f;
            o.z = (4 * cg + 2 < C) ? __ldg(in + (int64_t)(4 * cg + 2) * n + v) : 0.f;
            o.w = (4 * cg + 3 < C) ? __ldg(in + (int64_t)(4 * cg + 3) * n + v) : 0.f;
            out[v * vs + cg * gs] = o;
        }
}

// CGT > 0: the number of 4-channel groups is a compile-time constant (3 for the 12 MIND channels, 4 for 16 labels) and the
// group loop is unrolled, so that all 8 * CGT corner gathers of a voxel are in flight together (the kernel is bound by the
// latency of those L2 hits); CGT = 0: any channel count, one group at a time.
template <int CGT, int MINB>
__global__ void __launch_bounds__(128, MINB) k_adam_grad(const float* __restrict__ u, const float4* __restrict__ fixcl,
                                                   const float4* __restrict__ movcl, int CGrt, AdamGeom g,
                                                   const float* __restrict__ g0x, const float* __restrict__ g0y,
                                                   const float* __restrict__ g0z, float* __restrict__ gu) {
    ca_grid_dep_launch();      // programmatic dependent launch: the next kernel of the iteration may be scheduled ...
    ca_grid_dep_wait();        // ... and this one reads nothing before its predecessor (u = B^3 p) has completed
    const int CG = CGT > 0 ? CGT : CGrt;
    const unsigned long long pol_first = ca_l2_policy_evict_first(), pol_last = ca_l2_policy_evict_last();
    (void)pol_first; (void)pol_last;
    const int ha = g.ha, wa = g.wa, da = g.da;
    const int n = g.hl * wa * da;                   // field arrays: window-local indexing
    const int z = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int)(blockIdx.y * blockDim.y + threadIdx.y);
    const int xr = (int)(blockIdx.z * blockDim.z + threadIdx.z), xl = g.flip ? (int)(gridDim.z * blockDim.z) - 1 - xr : xr, x = xl + g.lo;
    if (z >= da || y >= wa || xl >= g.hl) return;
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
    // (an explicit register double buffer of the next group's gather was measured slower than one more resident CTA)
    float4 val[8], f;
#pragma unroll
    for (int cg = 0; cg < (CGT > 0 ? CGT : CG); ++cg) {
#pragma unroll
        for (int t = 0; t < 8; ++t) val[t] = inb[t] ? CA_GRAD_LD_MOV(movcl + off[t] + cg) : make_float4(0.f, 0.f, 0.f, 0.f);
        f = CA_GRAD_LD_FIX(fixcl + of * CG + cg);
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

// Second form of the gradient kernel (default): the same values with about a third fewer instructions -- the first form
// executes 1 040 instructions per voxel, 45 % of them before the channel loop, at 52 % of the issue slots
// (profiles/r2_ncu_k_adam_grad.txt).  (a) Corners outside the volume are not predicated: their coordinates are clamped, so every
// gather is a plain load, and their WEIGHT is zero -- `smp + val * 0` is `smp`, bit for bit, which is torch skipping the corner
// (GridSampler.h within_bounds); (b) the dot products residual . corner accumulate over the channel groups and the twelve corner
// differences are formed once after the loop (the first form did that per group; the order of these additions is not the
// reference's in either form -- autograd's is not restatable -- and the result stays inside the same tolerance);
// (c) the regulariser reads its neighbours at clamped offsets instead of under predicates (`uc - uc` adds +0);
// (d) x / 2 is x * 0.5 (exact).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_adam_grad2(const float* __restrict__ u, const float4* __restrict__ fixcl,
                                                    const float4* __restrict__ movcl, int CG, AdamGeom g,
                                                    const float* __restrict__ g0x, const float* __restrict__ g0y,
                                                    const float* __restrict__ g0z, float* __restrict__ gu) {
    ca_grid_dep_launch();
    ca_grid_dep_wait();
    const int ha = g.ha, wa = g.wa, da = g.da;
    const int n = g.hl * wa * da;
    const int z = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const int y = (int)(blockIdx.y * blockDim.y + threadIdx.y);
    const int xr = (int)(blockIdx.z * blockDim.z + threadIdx.z), xl = g.flip ? (int)(gridDim.z * blockDim.z) - 1 - xr : xr, x = xl + g.lo;
    if (z >= da || y >= wa || xl >= g.hl) return;
    const int o = (xl * wa + y) * da + z;
    const int of = (x * wa + y) * da + z;
    const float uc[3] = {u[o], u[n + o], u[2 * n + o]};
    float greg[3];
    {
        // a window edge that is not a face of the grid has no neighbour plane: that row of the window is margin the caller
        // discards (slab.py); a missing neighbour is read as the voxel itself and contributes uc - uc = +0
        const int dlo[3] = {xl > 0 ? -wa * da : 0, y > 0 ? -da : 0, z > 0 ? -1 : 0};
        const int dhi[3] = {xl < g.hl - 1 ? wa * da : 0, y < wa - 1 ? da : 0, z < da - 1 ? 1 : 0};
        float lo[3][3], hi[3][3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch)
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                lo[ch][ax] = u[ch * n + o + dlo[ax]];
                hi[ch][ax] = u[ch * n + o + dhi[ax]];
            }
        // (a face of the GRID has a one-sided difference: there the window ends too, and the missing side adds +0)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float gr = 0.f;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) gr += g.reg[ax] * ((uc[ch] - lo[ch][ax]) + (uc[ch] - hi[ch][ax]));
            greg[ch] = gr;
        }
    }
    // grid_disp = grid0 + (u/scale).flip(1)   (MIND.py:171), then GridSampler.h:27-36, every op rounded as torch does
    const float cx = __fadd_rn(g0x[z], ca_div(uc[2], g.sc[2]));
    const float cy = __fadd_rn(g0y[y], ca_div(uc[1], g.sc[1]));
    const float cz = __fadd_rn(g0z[x], ca_div(uc[0], g.sc[0]));
    const float ix = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(cx, 1.f), (float)da), 1.f), 0.5f);
    const float iy = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(cy, 1.f), (float)wa), 1.f), 0.5f);
    const float iz = __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(cz, 1.f), (float)ha), 1.f), 0.5f);
    const float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
    const float tx = __fsub_rn(ix, fx), ty = __fsub_rn(iy, fy), tz = __fsub_rn(iz, fz);
    const float sx = __fsub_rn(__fadd_rn(fx, 1.f), ix), sy = __fsub_rn(__fadd_rn(fy, 1.f), iy), sz = __fsub_rn(__fadd_rn(fz, 1.f), iz);
    const int x0 = (int)fminf(fmaxf(fx, -2.f), 1.0e9f), y0 = (int)fminf(fmaxf(fy, -2.f), 1.0e9f), z0 = (int)fminf(fmaxf(fz, -2.f), 1.0e9f);
    // per axis: is the low / high corner inside, and its clamped coordinate
    const bool vx[2] = {x0 >= 0 && x0 < da, x0 + 1 >= 0 && x0 + 1 < da};
    const bool vy[2] = {y0 >= 0 && y0 < wa, y0 + 1 >= 0 && y0 + 1 < wa};
    const bool vz[2] = {z0 >= 0 && z0 < ha, z0 + 1 >= 0 && z0 + 1 < ha};
    const int xc[2] = {ca_min(ca_max(x0, 0), da - 1) * g.vs, ca_min(ca_max(x0 + 1, 0), da - 1) * g.vs};
    const int yc[2] = {ca_min(ca_max(y0, 0), wa - 1), ca_min(ca_max(y0 + 1, 0), wa - 1)};
    const int zc[2] = {ca_min(ca_max(z0, 0), ha - 1), ca_min(ca_max(z0 + 1, 0), ha - 1)};
    float wgt[8];
    wgt[0] = __fmul_rn(__fmul_rn(sx, sy), sz); wgt[1] = __fmul_rn(__fmul_rn(tx, sy), sz);
    wgt[2] = __fmul_rn(__fmul_rn(sx, ty), sz); wgt[3] = __fmul_rn(__fmul_rn(tx, ty), sz);
    wgt[4] = __fmul_rn(__fmul_rn(sx, sy), tz); wgt[5] = __fmul_rn(__fmul_rn(tx, sy), tz);
    wgt[6] = __fmul_rn(__fmul_rn(sx, ty), tz); wgt[7] = __fmul_rn(__fmul_rn(tx, ty), tz);
    int off[8];
    unsigned inmask = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const bool in = vx[t & 1] && vy[(t >> 1) & 1] && vz[(t >> 2) & 1];
        if (!in) wgt[t] = 0.f;
        inmask |= in ? 1u << t : 0u;
        off[t] = (zc[(t >> 2) & 1] * wa + yc[(t >> 1) & 1]) * da * g.vs + xc[t & 1];
    }
    float dot[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) dot[t] = 0.f;
    const float4* fp = fixcl + (int64_t)of * g.vs;
    for (int cg = 0; cg < CG; ++cg, movcl += g.gs, fp += g.gs) {
        float4 val[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) val[t] = __ldg(movcl + off[t]);
        const float4 f = __ldg(fp);
        float4 smp = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 8; ++t) {      // torch's corner order (tnw, tne, tsw, tse, bnw, ...), products and sums rounded separately
            smp.x = __fadd_rn(smp.x, __fmul_rn(val[t].x, wgt[t]));
            smp.y = __fadd_rn(smp.y, __fmul_rn(val[t].y, wgt[t]));
            smp.z = __fadd_rn(smp.z, __fmul_rn(val[t].z, wgt[t]));
            smp.w = __fadd_rn(smp.w, __fmul_rn(val[t].w, wgt[t]));
        }
        const float rx = smp.x - f.x, ry = smp.y - f.y, rz = smp.z - f.z, rw = smp.w - f.w;
#pragma unroll
        for (int t = 0; t < 8; ++t) dot[t] += rx * val[t].x + ry * val[t].y + rz * val[t].z + rw * val[t].w;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t)
        if (!(inmask >> t & 1u)) dot[t] = 0.f;          // a corner outside the volume carries no gradient
    // d(w_t)/d(ix) = +-(Y_t Z_t) etc.: sum_t val_t dw_t = sum over corner pairs of (val_hi - val_lo) * (product of the other two)
    const float gix = (sy * sz) * (dot[1] - dot[0]) + (ty * sz) * (dot[3] - dot[2]) + (sy * tz) * (dot[5] - dot[4]) + (ty * tz) * (dot[7] - dot[6]);
    const float giy = (sx * sz) * (dot[2] - dot[0]) + (tx * sz) * (dot[3] - dot[1]) + (sx * tz) * (dot[6] - dot[4]) + (tx * tz) * (dot[7] - dot[5]);
    const float giz = (sx * sy) * (dot[4] - dot[0]) + (tx * sy) * (dot[5] - dot[1]) + (sx * ty) * (dot[6] - dot[2]) + (tx * ty) * (dot[7] - dot[3]);
    gu[2 * n + o] = g.gscale * gix * g.gmul[2] + greg[2];
    gu[n + o] = g.gscale * giy * g.gmul[1] + greg[1];
    gu[o] = g.gscale * giz * g.gmul[0] + greg[0];
}

#ifndef CA_GRAD_MINB
#define CA_GRAD_MINB 5
#endif
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

// ---- B^3 = three F.avg_pool3d(x, 3, stride=1, padding=1) passes (MIND.py:166), fused ---------------------------------
// box3 with zero padding is separable per axis, and so is the truncation of each intermediate to the volume, hence
//     B^3 = T_H T_W T_D,   T = (1-D box3 with zero padding)^3 along one axis, / 27.
// Two kernels instead of three full 27-tap passes: k_box3_h applies T_H (a thread owns a column segment and slides
// three cascaded 3-tap sums along it in registers), k_box3_p applies T_W T_D to one plane held in shared memory and,
// in UPDATE mode, feeds the finished value dL/dp straight into the Adam step of (p, m, v).
// The additions are associated differently from torch's ordered 27-tap sums: u agrees with the reference's to a few
// ulp (tests: 2e-6 relative), which the stage's own tolerance (Adam amplifies 1-ulp feature noise) absorbs.
struct Box3Args {
    const float* in;
    float* out;
    int h, w, d, nq;
    int Ty, ntiles, Ly;           // k_box3_p: output rows per CTA, row tiles per plane, rows per thread
    float *p, *m, *v;
    AdamScalars s;
    int flip;                     // 1: segments / planes and channels are walked in descending order (see AdamGeom::flip)
};

static __device__ __forceinline__ float4 f4add3(const float4& a, const float4& b, const float4& c) {
    return make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
}
static __device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// T_H: grid = (ceil(w*nq / 128), H segments, 3 channels); a thread owns (row y, quad q) and the planes [xs, xe)
template <int LX>
__global__ void __launch_bounds__(128) k_box3_h(Box3Args a) {
    ca_grid_dep_launch();
    ca_grid_dep_wait();
    const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (t >= a.w * a.nq) return;
    const int by = a.flip ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y, bz = a.flip ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    const int xs = by * LX, xe = ca_min(a.h, xs + LX);
    const int64_t plane = (int64_t)a.w * a.d;
    const float* ip = a.in + (int64_t)bz * a.h * plane + (int64_t)t * 4;
    float* op = a.out + (int64_t)bz * a.h * plane + (int64_t)t * 4;
    float4 v[LX + 6];
#pragma unroll
    for (int i = 0; i < LX + 6; ++i) {           // all loads of the segment go out together
        const int X = xs - 3 + i;
        v[i] = (X >= 0 && X < a.h && X < xe + 3) ? __ldg(reinterpret_cast<const float4*>(ip + (int64_t)X * plane)) : f4zero();
    }
    float4 i2 = f4zero(), i1 = f4zero(), a2 = f4zero(), a1 = f4zero(), b2 = f4zero(), b1 = f4zero();
    const float c27 = 1.0f / 27.0f;
#pragma unroll
    for (int i = 0; i < LX + 6; ++i) {
        const int X = xs - 3 + i;                // plane X arrives: a[X-1], b[X-2], c[X-3] complete
        float4 an = f4add3(i2, i1, v[i]);
        if (X - 1 < 0 || X - 1 >= a.h) an = f4zero();
        float4 bn = f4add3(a2, a1, an);
        if (X - 2 < 0 || X - 2 >= a.h) bn = f4zero();
        const float4 cn = f4add3(b2, b1, bn);
        if (X - 3 >= xs && X - 3 < xe)
            *reinterpret_cast<float4*>(op + (int64_t)(X - 3) * plane) = make_float4(cn.x * c27, cn.y * c27, cn.z * c27, cn.w * c27);
        i2 = i1; i1 = v[i]; a2 = a1; a1 = an; b2 = b1; b1 = bn;
    }
}

// T_W T_D on one plane: grid = (row tiles, h, 3 channels).  Shared memory: two buffers of (Ty + 6) rows x d floats.
template <bool UPDATE>
__global__ void __launch_bounds__(256) k_box3_p(Box3Args a) {
    ca_grid_dep_launch();
    ca_grid_dep_wait();
    CA_DYN_SMEM(float, sm);
    const int tid = (int)threadIdx.x, nthr = (int)blockDim.x;
    const int ty0 = (int)blockIdx.x * a.Ty;
    const int nrows = ca_min(a.Ty, a.w - ty0) + 6;              // rows ty0-3 .. ty0+Ty+2 of the plane
    const int nq = a.nq, d = a.d;
    float* A = sm;
    float* B = sm + (size_t)(a.Ty + 6) * d;
    const int by = a.flip ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y, bz = a.flip ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    const int64_t plane_off = ((int64_t)bz * a.h + by) * a.w * d;
    for (int i = tid; i < nrows * nq; i += nthr) {
        const int r = i / nq, q = i - r * nq, y = ty0 - 3 + r;
        *reinterpret_cast<float4*>(A + r * d + 4 * q) =
            (y >= 0 && y < a.w) ? __ldg(reinterpret_cast<const float4*>(a.in + plane_off + (int64_t)y * d + 4 * q)) : f4zero();
    }
    __syncthreads();
    // T_D: the quad with its neighbours; positions z0-4 .. z0+7 -> r[0..11]
    for (int i = tid; i < nrows * nq; i += nthr) {
        const int rr = i / nq, q = i - rr * nq;
        const float* row = A + rr * d;
        const float4 lo = q > 0 ? *reinterpret_cast<const float4*>(row + 4 * q - 4) : f4zero();
        const float4 mi = *reinterpret_cast<const float4*>(row + 4 * q);
        const float4 hi = q < nq - 1 ? *reinterpret_cast<const float4*>(row + 4 * q + 4) : f4zero();
        const float r[12] = {lo.x, lo.y, lo.z, lo.w, mi.x, mi.y, mi.z, mi.w, hi.x, hi.y, hi.z, hi.w};
        float s1[12], s2[12];
#pragma unroll
        for (int j = 2; j <= 9; ++j) s1[j] = (r[j - 1] + r[j]) + r[j + 1];
        if (q == 0) s1[2] = s1[3] = 0.f;                        // positions -2, -1 lie outside the volume
        if (q == nq - 1) s1[8] = s1[9] = 0.f;                   // positions d, d+1
#pragma unroll
        for (int j = 3; j <= 8; ++j) s2[j] = (s1[j - 1] + s1[j]) + s1[j + 1];
        if (q == 0) s2[3] = 0.f;
        if (q == nq - 1) s2[8] = 0.f;
        float4 o;
        o.x = (s2[3] + s2[4]) + s2[5]; o.y = (s2[4] + s2[5]) + s2[6]; o.z = (s2[5] + s2[6]) + s2[7]; o.w = (s2[6] + s2[7]) + s2[8];
        *reinterpret_cast<float4*>(B + rr * d + 4 * q) = o;
    }
    __syncthreads();
    // T_W: a thread owns quad q and the output rows [y0, y0+Ly) of the tile and slides along W
    const float c729 = 1.0f / 729.0f;
    const int nchunks = (a.Ty + a.Ly - 1) / a.Ly;
    for (int i = tid; i < nchunks * nq; i += nthr) {
        const int ch = i / nq, q = i - ch * nq;
        const int y0 = ty0 + ch * a.Ly, y1 = ca_min(ca_min(y0 + a.Ly, ty0 + a.Ty), a.w);
        float4 i2 = f4zero(), i1 = f4zero(), a2 = f4zero(), a1 = f4zero(), b2 = f4zero(), b1 = f4zero();
        // UPDATE: the (p, m, v) rows are requested three rows ahead of their use (a queue of three register sets,
        // statically indexed: the row loop is unrolled by three)
        float4 pq[3], mq[3], vq[3];
        const int64_t col = plane_off + 4 * q;
        auto step = [&](int Y, int slot) {                      // row Y arrives: a[Y-1], b[Y-2], c[Y-3] complete
            const float4 vin = (Y >= 0 && Y < a.w) ? *reinterpret_cast<const float4*>(B + (Y - ty0 + 3) * d + 4 * q) : f4zero();
            float4 an = f4add3(i2, i1, vin);
            if (Y - 1 < 0 || Y - 1 >= a.w) an = f4zero();
            float4 bn = f4add3(a2, a1, an);
            if (Y - 2 < 0 || Y - 2 >= a.w) bn = f4zero();
            const float4 cn = f4add3(b2, b1, bn);
            if (Y - 3 >= y0) {
                const int64_t off = col + (int64_t)(Y - 3) * d;
                float4 r = make_float4(cn.x * c729, cn.y * c729, cn.z * c729, cn.w * c729);
                if (UPDATE) {
                    float4 pp = pq[slot], mm = mq[slot], vv = vq[slot];
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
            if (UPDATE && Y >= y0 && Y < y1) {                  // consumed at step Y + 3, same slot
                const int64_t off = col + (int64_t)Y * d;
                pq[slot] = *reinterpret_cast<const float4*>(a.p + off);
                mq[slot] = *reinterpret_cast<const float4*>(a.m + off);
                vq[slot] = *reinterpret_cast<const float4*>(a.v + off);
            }
            i2 = i1; i1 = vin; a2 = a1; a1 = an; b2 = b1; b1 = bn;
        };
        for (int Y = y0 - 3; Y < y1 + 3; Y += 3) {
            step(Y, 0);
            if (Y + 1 < y1 + 3) step(Y + 1, 1);
            if (Y + 2 < y1 + 3) step(Y + 2, 2);
        }
    }
}

// ---- the search's GaussianSmoothing (convexAdam_hyper_util.py:425-473): per axis, replicate padding by N/2 and a valid
// cross-correlation with N normalised Gaussian weights, out[i] = sum_t w[t] in[clamp(i + t - N/2)].  adjoint != 0 applies the
// transpose of that matrix (the backward of the padded convolution: what the clamped taps read is added back onto the border
// voxel).  One thread per output element, taps in ascending order.
struct Conv1dArgs { float w[32]; int N; };
__global__ void __launch_bounds__(256) k_conv1d_rep(const float* __restrict__ in, float* __restrict__ out, int nc, int h, int w, int d,
                                                    int axis, Conv1dArgs a, int adjoint) {
    const int n = h * w * d, total = nc * n;
    const int len = axis == 0 ? h : (axis == 1 ? w : d), stride = axis == 0 ? w * d : (axis == 1 ? d : 1), r = a.N / 2;
    for (int o = (int)(blockIdx.x * blockDim.x + threadIdx.x); o < total; o += (int)(gridDim.x * blockDim.x)) {
        const int v = o % n;
        const int j = axis == 0 ? v / (w * d) : (axis == 1 ? (v / d) % w : v % d);
        const float* line = in + (o - j * stride);
        float acc = 0.f;
        if (!adjoint) {
            for (int t = 0; t < a.N; ++t) acc += a.w[t] * line[ca_min(ca_max(j + t - r, 0), len - 1) * stride];
        } else {
            // every (i, t) with clamp(i + t - r) == j, i ascending then t ascending
            for (int i = ca_max(j - r, 0); i <= ca_min(j + r, len - 1); ++i) {
                const float g = line[i * stride];
                for (int t = 0; t < a.N; ++t) {
                    const int q = i + t - r;
                    if ((q == j) || (j == 0 && q < 0) || (j == len - 1 && q > len - 1)) acc += a.w[t] * g;
                }
            }
        }
        out[o] = acc;
    }
}

static inline unsigned adam_grid(int64_t n, int block) {
    int64_t b = ca_divup(n, block);
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

#define CA_BOX3_LX 6
#define CA_BOX3_SMEM (96 * 1024)

// tuning knobs, read once per process (defaults = the measured best)
static int adam_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e && *e ? atoi(e) : dflt;
}
struct AdamKnobs { int grad, flip, bx, by, bz, ty, lx, layout, minb; };
static const AdamKnobs& adam_knobs() {
    static const AdamKnobs k = {
        adam_env("CA_ADAM_GRAD", 2),     // 1: first form of the gradient kernel, 2: k_adam_grad2
        adam_env("CA_ADAM_FLIP", 1),     // bit 0: gradient kernel, bit 1: box kernels walk the planes downwards on odd iterations
        adam_env("CA_ADAM_BX", 32), adam_env("CA_ADAM_BY", 4), adam_env("CA_ADAM_BZ", 1),    // voxels per CTA of the gradient kernel (product 128)
        adam_env("CA_BOX3_TY", 0),       // rows per CTA of k_box3_p (0: as many as fit 96 KB of shared memory)
        adam_env("CA_BOX3_LX", CA_BOX3_LX),   // planes per thread of k_box3_h: 6, 8, 12 or 16
        adam_env("CA_ADAM_LAYOUT", 1),   // feature groups 0: (n, CG) voxel-major, 1: (CG, n) group-planar (k_adam_grad2 only)
        adam_env("CA_ADAM_MINB", CA_GRAD_MINB),   // resident CTAs per SM the gradient kernel is compiled for: 4, 5, 6 or 8
    };
    return k;
}

// rows per CTA of k_box3_p such that both shared-memory buffers fit; 0: the plane rows are too long
static int box3_rows_per_tile(int wa, int da) {
    int maxrows = (int)(CA_BOX3_SMEM / ((size_t)2 * da * 4)) - 6;
    if (maxrows < 1) return 0;
    if (adam_knobs().ty > 0 && adam_knobs().ty < maxrows) maxrows = adam_knobs().ty;
    const int nt = (wa + maxrows - 1) / maxrows;
    return (wa + nt - 1) / nt;
}

// out = B^3 in on a (3, h, w, d) field, through the scratch field tmp; with `upd` the result drives the Adam step
// (and is stored only if out != null)
static int box3_cubed(const float* in, float* tmp, float* out, int ha, int wa, int da, const Box3Args* upd, int flip, cudaStream_t st) {
    Box3Args a;
    if (upd) a = *upd;
    a.flip = flip;
    a.h = ha; a.w = wa; a.d = da; a.nq = da / 4;
    a.in = in; a.out = tmp;
    {
        const int lx = adam_knobs().lx;
        const dim3 grid((unsigned)ca_divup((int64_t)wa * a.nq, 128), (unsigned)ca_divup(ha, lx), 3u);
        switch (lx) {
            case 6: CA_LAUNCH_PDL(k_box3_h<6>, grid, dim3(128), 0, st, a); break;
            case 8: CA_LAUNCH_PDL(k_box3_h<8>, grid, dim3(128), 0, st, a); break;
            case 12: CA_LAUNCH_PDL(k_box3_h<12>, grid, dim3(128), 0, st, a); break;
            case 16: CA_LAUNCH_PDL(k_box3_h<16>, grid, dim3(128), 0, st, a); break;
            default: ca_set_error("adam_run: CA_BOX3_LX must be 6, 8, 12 or 16"); return CA_ERR_BAD_ARGUMENT;
        }
        CA_CHECK_LAUNCH("box3_h");
    }
    a.in = tmp; a.out = out;
    a.Ty = box3_rows_per_tile(wa, da);
    a.ntiles = (wa + a.Ty - 1) / a.Ty;
    const int cols = 256 / a.nq > 0 ? 256 / a.nq : 1;       // row chunks that keep 256 threads busy
    a.Ly = (a.Ty + cols - 1) / cols;
    if (a.Ly < 4) a.Ly = ca_min(4, a.Ty);
    const size_t smem = (size_t)2 * (a.Ty + 6) * da * 4;
    const dim3 grid((unsigned)a.ntiles, (unsigned)ha, 3u);
    if (upd) { auto k = k_box3_p<true>; CA_SMEM_OPTIN(k, CA_BOX3_SMEM + 6 * 4096); CA_LAUNCH_PDL(k, grid, dim3(256), smem, st, a); }
    else { auto k = k_box3_p<false>; CA_SMEM_OPTIN(k, CA_BOX3_SMEM + 6 * 4096); CA_LAUNCH_PDL(k, grid, dim3(256), smem, st, a); }
    CA_CHECK_LAUNCH("box3_p");
    return CA_OK;
}

extern "C" size_t ca_adam_window_workspace_bytes(int C, int ha, int wa, int da, int lo, int hi) {
    if (C <= 0 || ha <= 0 || wa <= 0 || da <= 0 || lo < 0 || hi > ha || lo >= hi) return 0;
    const size_t n = (size_t)ha * wa * da, nw = (size_t)(hi - lo) * wa * da;
    return nw * 3 * 4 * 3 + 2 * n * (size_t)((C + 3) & ~3) * 4 + 1024 + ((size_t)da + 64) * 4;
}
extern "C" size_t ca_adam_workspace_bytes(int C, int ha, int wa, int da) { return ca_adam_window_workspace_bytes(C, ha, wa, da, 0, ha); }

// widths / nwidths: nullptr / 0 = the reference's three 3-wide pools (B^3, MIND.py:166); otherwise the smoother is the sequence
// AvgPool3d(w_i, stride 1, padding (w_i - 1) / 2) of the self-configuring search (kovesi_spline, convexAdam_hyper_util.py:475-488;
// used at adam_run_withconfig_shiftSpline.py:217).  Each pool is symmetric (zero padding, divisor w^3 everywhere), so the
// backward of the sequence is the same pools in reverse order.
static int adam_run_impl(const float* fix, const float* mov, int C, int ha, int wa, int da, int lo, int hi,
                                  float lambda_weight, float* p, float* m, float* v, int first_step, int niter, const float* g0x,
                                  const float* g0y, const float* g0z, float* u_out, float* g_out, void* workspace,
                                  size_t workspace_bytes, int features_ready, const int* widths, int nwidths, const float* gauss, int ngauss, ca_stream_t stream) {
    CA_CHECK_ARG(fix && mov && p && m && v && g0x && g0y && g0z && u_out && workspace, "adam_run: null pointer");
    CA_CHECK_ARG(C > 0 && C <= 128 && ha > 1 && wa > 1 && da > 1, "adam_run: bad dims (1..128 channels, each Adam-grid axis >= 2)");
    CA_CHECK_ARG(lo >= 0 && hi <= ha && lo < hi, "adam_run: bad window [%d, %d) of %d planes", lo, hi, ha);
    CA_CHECK_ARG(first_step >= 1 && niter >= 0, "adam_run: bad step range");
    // the kernels index the planar (3, n) fields and the (n, CG) feature groups with 32-bit integers
    CA_CHECK_ARG((int64_t)ha * wa * da * ((C + 3) / 4) < (1ll << 31) && (int64_t)3 * ha * wa * da < (1ll << 31) && wa < 65536 && ha < 65536,
                 "adam_run: grid too large");
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
    // the fused B^3 kernels need float4 rows that fit shared memory; any other shape takes three generic avg_pool3d passes
    const bool fast = nwidths == 0 && ngauss == 0 && (da & 3) == 0 && box3_rows_per_tile(wa, da) > 0 &&
                      (((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)u_out | (uintptr_t)g_out) & 15) == 0;
    (void)zero_row;
    auto box = [&](const float* in, float* out) -> int { return ca_launch_avgpool3d(in, 3, hl, wa, da, 3, 1, 1, out, st); };
    // the Gaussian smoother (forward: axes 0, 1, 2; backward: the transposed filters along 2, 1, 0), src -> dst through t0, t1
    Conv1dArgs cw;
    cw.N = ngauss;
    for (int i = 0; i < 32; ++i) cw.w[i] = (gauss && i < ngauss) ? gauss[i] : 0.f;
    auto gaussian = [&](const float* src, float* dst, bool adjoint) -> int {
        const unsigned gb = adam_grid(3 * nw, 256);
        const int ax[3] = {adjoint ? 2 : 0, 1, adjoint ? 0 : 2};
        CA_LAUNCH(k_conv1d_rep, dim3(gb), dim3(256), 0, st, src, t0, 3, hl, wa, da, ax[0], cw, (int)adjoint);
        CA_LAUNCH(k_conv1d_rep, dim3(gb), dim3(256), 0, st, (const float*)t0, t1, 3, hl, wa, da, ax[1], cw, (int)adjoint);
        CA_LAUNCH(k_conv1d_rep, dim3(gb), dim3(256), 0, st, (const float*)t1, dst, 3, hl, wa, da, ax[2], cw, (int)adjoint);
        CA_CHECK_LAUNCH("adam/gaussian");
        return CA_OK;
    };
    // the pool sequence src -> dst through the two scratch fields (forward: ascending, backward: descending order)
    auto pools = [&](const float* src, float* dst, bool reverse) -> int {
        const float* cur = src;
        for (int i = 0; i < nwidths; ++i) {
            const int w = widths[reverse ? nwidths - 1 - i : i];
            float* out = (i == nwidths - 1) ? dst : (cur == t0 ? t1 : t0);
            const int rc = ca_launch_avgpool3d(cur, 3, hl, wa, da, w, 1, (w - 1) / 2, out, st);
            if (rc != CA_OK) return rc;
            cur = out;
        }
        return CA_OK;
    };
    const AdamKnobs& kn = adam_knobs();
    CA_CHECK_ARG(kn.bx > 0 && kn.by > 0 && kn.bz > 0 && kn.bx * kn.by * kn.bz == 128, "adam_run: CA_ADAM_BX * BY * BZ must be 128");
    const bool planar_groups = kn.layout == 1 && kn.grad != 1;
    const int fvs = planar_groups ? 1 : CG, fgs = planar_groups ? (int)n : 1;
    if (niter > 0 && !features_ready) {
        const unsigned gb = adam_grid(n, 256);
        CA_LAUNCH(k_planar_to_cl, dim3(gb), dim3(256), 0, st, fix, C, CG, n, (int64_t)fvs, (int64_t)fgs, fixcl);
        CA_LAUNCH(k_planar_to_cl, dim3(gb), dim3(256), 0, st, mov, C, CG, n, (int64_t)fvs, (int64_t)fgs, movcl);
        CA_CHECK_LAUNCH("adam/planar_to_cl");
    }
    AdamGeom g;
    g.flip = 0; g.vs = fvs; g.gs = fgs;
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
        // odd steps walk the grid downwards: what the previous iteration touched last is what this one touches first, so the
        // part of the features + fields (165 MB at config 2) that the 126 MB L2 still holds is hit instead of cycled out
        const int flip_box = (kn.flip & 2) ? (step & 1) : 0;
        g.flip = (kn.flip & 1) ? (step & 1) : 0;
        if (fast) {
            if ((rc = box3_cubed(p, t0, u_out, hl, wa, da, nullptr, flip_box, st)) != CA_OK) return rc;
        } else if (nwidths > 0) {
            if ((rc = pools(p, u_out, false)) != CA_OK) return rc;
        } else if (ngauss > 0) {
            if ((rc = gaussian(p, u_out, false)) != CA_OK) return rc;
        } else {
            if ((rc = box(p, t0)) != CA_OK) return rc;
            if ((rc = box(t0, t1)) != CA_OK) return rc;
            if ((rc = box(t1, u_out)) != CA_OK) return rc;
        }
        const int bx = kn.bx, by = kn.by, bz = kn.bz;      // 32 x 4 voxels per CTA: neighbouring rows share corner gathers in L1
        void* tmg = ca_timing_begin(CA_T_ADAM_GRAD, st);
        const dim3 ggrid((unsigned)((da + bx - 1) / bx), (unsigned)((wa + by - 1) / by), (unsigned)((hl + bz - 1) / bz)), gblock(bx, by, bz);
        // (compile-time group counts with the loop unrolled -- all 24 gathers of a voxel in flight -- were measured slower: 9.5-10.6
        // against 8.9 ms per 100 iterations; the kernel wants resident warps more than loads per warp)
        if (kn.grad == 1)
            CA_LAUNCH_PDL((k_adam_grad<0, CA_GRAD_MINB>), ggrid, gblock, 0, st, (const float*)u_out, (const float4*)fixcl, (const float4*)movcl, CG, g, g0x, g0y, g0z, gu);
        else if (kn.minb == 4)
            CA_LAUNCH_PDL((k_adam_grad2<4>), ggrid, gblock, 0, st, (const float*)u_out, (const float4*)fixcl, (const float4*)movcl, CG, g, g0x, g0y, g0z, gu);
        else if (kn.minb == 6)
            CA_LAUNCH_PDL((k_adam_grad2<6>), ggrid, gblock, 0, st, (const float*)u_out, (const float4*)fixcl, (const float4*)movcl, CG, g, g0x, g0y, g0z, gu);
        else if (kn.minb == 8)
            CA_LAUNCH_PDL((k_adam_grad2<8>), ggrid, gblock, 0, st, (const float*)u_out, (const float4*)fixcl, (const float4*)movcl, CG, g, g0x, g0y, g0z, gu);
        else
            CA_LAUNCH_PDL((k_adam_grad2<CA_GRAD_MINB>), ggrid, gblock, 0, st, (const float*)u_out, (const float4*)fixcl, (const float4*)movcl, CG, g, g0x, g0y, g0z, gu);
        ca_timing_end(tmg, st);
        CA_CHECK_LAUNCH("adam_grad");
        if (!fast && nwidths == 0 && ngauss == 0) {
            if ((rc = box(gu, t0)) != CA_OK) return rc;
            if ((rc = box(t0, t1)) != CA_OK) return rc;
        }
        Box3Args upd;
        const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
        upd.p = p; upd.m = m; upd.v = v;
        upd.s.w_lerp = (float)(1.0 - b1);
        upd.s.beta2 = (float)b2;
        upd.s.one_m_beta2 = (float)(1.0 - b2);
        upd.s.bc2_sqrt = (float)sqrt(bc2);
        upd.s.eps = (float)eps;
        upd.s.neg_step_size = (float)(-(lr / bc1));
        if (fast) {
            // B^3 g_u fused with the Adam step; dL/dp is stored only when the caller asks for it
            if ((rc = box3_cubed(gu, t0, g_out, hl, wa, da, &upd, flip_box, st)) != CA_OK) return rc;
        } else {
            // (with a pool list dL/dp may overwrite g_u itself: the first pool has read it by then -- unless it is the only one)
            float* gp = g_out ? g_out : ((nwidths > 1 || ngauss > 0) ? gu : t0);
            if (nwidths > 0) { if ((rc = pools(gu, gp, true)) != CA_OK) return rc; }
            else if (ngauss > 0) { if ((rc = gaussian(gu, gp, true)) != CA_OK) return rc; }
            else if ((rc = box(t1, gp)) != CA_OK) return rc;
            CA_LAUNCH(k_adam_update, dim3(adam_grid(3 * nw, 256)), dim3(256), 0, st, p, m, v, gp, 3 * nw, upd.s);
            CA_CHECK_LAUNCH("adam_update");
        }
    }
    return CA_OK;
}

extern "C" int ca_adam_run_window(const float* fix, const float* mov, int C, int ha, int wa, int da, int lo, int hi,
                                  float lambda_weight, float* p, float* m, float* v, int first_step, int niter, const float* g0x,
                                  const float* g0y, const float* g0z, float* u_out, float* g_out, void* workspace,
                                  size_t workspace_bytes, int features_ready, ca_stream_t stream) {
    return adam_run_impl(fix, mov, C, ha, wa, da, lo, hi, lambda_weight, p, m, v, first_step, niter, g0x, g0y, g0z, u_out, g_out,
                         workspace, workspace_bytes, features_ready, nullptr, 0, nullptr, 0, stream);
}

extern "C" int ca_adam_run_pools(const float* fix, const float* mov, int C, int ha, int wa, int da, float lambda_weight, float* p,
                                 float* m, float* v, int first_step, int niter, const float* g0x, const float* g0y,
                                 const float* g0z, const int* widths, int nwidths, float* u_out, float* g_out, void* workspace,
                                 size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(widths && nwidths >= 1 && nwidths <= 16, "adam_run_pools: 1..16 pool widths (host array)");
    for (int i = 0; i < nwidths; ++i) CA_CHECK_ARG(widths[i] >= 1 && (widths[i] & 1) == 1 && widths[i] <= 31, "adam_run_pools: pool widths must be odd, 1..31");
    return adam_run_impl(fix, mov, C, ha, wa, da, 0, ha, lambda_weight, p, m, v, first_step, niter, g0x, g0y, g0z, u_out, g_out,
                         workspace, workspace_bytes, 0, widths, nwidths, nullptr, 0, stream);
}

extern "C" int ca_adam_run_gauss(const float* fix, const float* mov, int C, int ha, int wa, int da, float lambda_weight, float* p,
                                 float* m, float* v, int first_step, int niter, const float* g0x, const float* g0y,
                                 const float* g0z, const float* weights, int nweights, float* u_out, float* g_out, void* workspace,
                                 size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(weights && nweights >= 1 && nweights <= 31 && (nweights & 1) == 1, "adam_run_gauss: an odd number (1..31) of filter weights (host array)");
    return adam_run_impl(fix, mov, C, ha, wa, da, 0, ha, lambda_weight, p, m, v, first_step, niter, g0x, g0y, g0z, u_out, g_out,
                         workspace, workspace_bytes, 0, nullptr, 0, weights, nweights, stream);
}

extern "C" int ca_adam_run(const float* fix, const float* mov, int C, int ha, int wa, int da, float lambda_weight, float* p,
                           float* m, float* v, int first_step, int niter, const float* g0x, const float* g0y,
                           const float* g0z, float* u_out, float* g_out, void* workspace, size_t workspace_bytes,
                           ca_stream_t stream) {
    return ca_adam_run_window(fix, mov, C, ha, wa, da, 0, ha, lambda_weight, p, m, v, first_step, niter, g0x, g0y, g0z, u_out,
                              g_out, workspace, workspace_bytes, 0, stream);
}
