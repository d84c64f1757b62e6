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
#include "ca_platform.h"

int ca_launch_avgpool3d(const float* in, int64_t nc, int T, int H, int W, int k, int stride, int pad, float* out,
                        cudaStream_t st);

struct AdamGeom {
    int ha, wa, da, C;
    float sc[3];      // (size-1)/2 per axis  (MIND.py:170)
    float gmul[3];    // d(position)/d(u) = (size/2) / sc
    float reg[3];     // 2*lambda / count per axis (MIND.py:167-169)
    float gscale;     // 2 * 12 / (C * N)  (MIND.py:176-177: mean over channels, *12, mean over voxels)
};

__global__ void __launch_bounds__(128) k_adam_grad(const float* __restrict__ u, const float* __restrict__ fix,
                                                   const float* __restrict__ mov, AdamGeom g,
                                                   const float* __restrict__ g0x, const float* __restrict__ g0y,
                                                   const float* __restrict__ g0z, float* __restrict__ gu) {
    const int ha = g.ha, wa = g.wa, da = g.da;
    const int64_t n = (int64_t)ha * wa * da;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(o % da), y = (int)((o / da) % wa), x = (int)(o / ((int64_t)da * wa));
        const float u0 = u[o], u1 = u[n + o], u2 = u[2 * n + o];
        // grid_disp = grid0 + (u/scale).flip(1)   (MIND.py:171), then GridSampler.h:27-36
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
        int64_t off[8];
        bool inb[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int xx = x0 + (t & 1), yy = y0 + ((t >> 1) & 1), zz = z0 + ((t >> 2) & 1);
            inb[t] = xx >= 0 && xx < da && yy >= 0 && yy < wa && zz >= 0 && zz < ha;
            off[t] = inb[t] ? ((int64_t)zz * wa + yy) * da + xx : 0;
        }
        float gix = 0.f, giy = 0.f, giz = 0.f;
        for (int c = 0; c < g.C; ++c) {
            const float* mp = mov + c * n;
            float val[8];
            float smp = 0.f;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                val[t] = inb[t] ? __ldg(mp + off[t]) : 0.f;
                if (inb[t]) smp = __fadd_rn(smp, __fmul_rn(val[t], wgt[t]));
            }
            const float r = __fsub_rn(smp, __ldg(fix + c * n + o));
            // d(sample)/d(ix,iy,iz) of the trilinear interpolant (zeros outside the volume)
            const float dx = sy * sz * (val[1] - val[0]) + ty * sz * (val[3] - val[2]) + sy * tz * (val[5] - val[4]) + ty * tz * (val[7] - val[6]);
            const float dy = sx * sz * (val[2] - val[0]) + tx * sz * (val[3] - val[1]) + sx * tz * (val[6] - val[4]) + tx * tz * (val[7] - val[5]);
            const float dz = sx * sy * (val[4] - val[0]) + tx * sy * (val[5] - val[1]) + sx * ty * (val[6] - val[2]) + tx * ty * (val[7] - val[3]);
            gix += r * dx; giy += r * dy; giz += r * dz;
        }
        float grad[3];
        grad[2] = g.gscale * gix * g.gmul[2];
        grad[1] = g.gscale * giy * g.gmul[1];
        grad[0] = g.gscale * giz * g.gmul[0];
        // diffusion regulariser: d/du_c(x) of lambda * mean_k (u[x+e_k]-u[x])^2, one-sided at the faces
        const int idx[3] = {x, y, z};
        const int dims[3] = {ha, wa, da};
        const int64_t strd[3] = {(int64_t)wa * da, (int64_t)da, 1};
        const float uc[3] = {u0, u1, u2};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            float gr = 0.f;
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
                float lap = 0.f;
                if (idx[ax] > 0) lap += uc[ch] - u[ch * n + o - strd[ax]];
                if (idx[ax] < dims[ax] - 1) lap += uc[ch] - u[ch * n + o + strd[ax]];
                gr += g.reg[ax] * lap;
            }
            gu[ch * n + o] = grad[ch] + gr;
        }
    }
}

struct AdamScalars { float w_lerp, beta2, one_m_beta2, bc2_sqrt, eps, neg_step_size; };

__global__ void __launch_bounds__(256) k_adam_update(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                     const float* __restrict__ g, int64_t n3, AdamScalars s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = __fadd_rn(m[i], __fmul_rn(s.w_lerp, __fsub_rn(gi, m[i])));                    // exp_avg.lerp_(grad, 1-beta1)
        const float vi = __fadd_rn(__fmul_rn(v[i], s.beta2), __fmul_rn(__fmul_rn(s.one_m_beta2, gi), gi));   // mul_(beta2).addcmul_
        const float denom = __fadd_rn(ca_div(__fsqrt_rn(vi), s.bc2_sqrt), s.eps);
        m[i] = mi;
        v[i] = vi;
        p[i] = __fadd_rn(p[i], __fmul_rn(s.neg_step_size, ca_div(mi, denom)));                      // addcdiv_(value=-step_size)
    }
}

static inline unsigned adam_grid(int64_t n, int block) {
    int64_t b = ca_divup(n, block);
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

extern "C" size_t ca_adam_workspace_bytes(int ha, int wa, int da) { return (size_t)ha * wa * da * 3 * 4 * 3 + 256; }

extern "C" int ca_adam_run(const float* fix, const float* mov, int C, int ha, int wa, int da, float lambda_weight, float* p,
                           float* m, float* v, int first_step, int niter, const float* g0x, const float* g0y,
                           const float* g0z, float* u_out, float* g_out, void* workspace, size_t workspace_bytes,
                           ca_stream_t stream) {
    CA_CHECK_ARG(fix && mov && p && m && v && g0x && g0y && g0z && u_out && workspace, "adam_run: null pointer");
    CA_CHECK_ARG(C > 0 && ha > 1 && wa > 1 && da > 1, "adam_run: bad dims (each Adam-grid axis must be >= 2)");
    CA_CHECK_ARG(first_step >= 1 && niter >= 0, "adam_run: bad step range");
    if (workspace_bytes < ca_adam_workspace_bytes(ha, wa, da)) {
        ca_set_error("adam_run: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)ha * wa * da;
    float* t0 = (float*)workspace;
    float* t1 = t0 + 3 * n;
    float* gu = t1 + 3 * n;
    AdamGeom g;
    g.ha = ha; g.wa = wa; g.da = da; g.C = C;
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
        if ((rc = ca_launch_avgpool3d(p, 3, ha, wa, da, 3, 1, 1, t0, st)) != CA_OK) return rc;
        if ((rc = ca_launch_avgpool3d(t0, 3, ha, wa, da, 3, 1, 1, t1, st)) != CA_OK) return rc;
        if ((rc = ca_launch_avgpool3d(t1, 3, ha, wa, da, 3, 1, 1, u_out, st)) != CA_OK) return rc;
        CA_LAUNCH(k_adam_grad, dim3(adam_grid(n, 128)), dim3(128), 0, st, u_out, fix, mov, g, g0x, g0y, g0z, gu);
        CA_CHECK_LAUNCH("adam_grad");
        if ((rc = ca_launch_avgpool3d(gu, 3, ha, wa, da, 3, 1, 1, t0, st)) != CA_OK) return rc;
        if ((rc = ca_launch_avgpool3d(t0, 3, ha, wa, da, 3, 1, 1, t1, st)) != CA_OK) return rc;
        float* gp = g_out ? g_out : t0;
        if ((rc = ca_launch_avgpool3d(t1, 3, ha, wa, da, 3, 1, 1, gp, st)) != CA_OK) return rc;
        AdamScalars s;
        const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
        s.w_lerp = (float)(1.0 - b1);
        s.beta2 = (float)b2;
        s.one_m_beta2 = (float)(1.0 - b2);
        s.bc2_sqrt = (float)sqrt(bc2);
        s.eps = (float)eps;
        s.neg_step_size = (float)(-(lr / bc1));
        CA_LAUNCH(k_adam_update, dim3(adam_grid(3 * n, 256)), dim3(256), 0, st, p, m, v, gp, 3 * n, s);
        CA_CHECK_LAUNCH("adam_update");
    }
    return CA_OK;
}
