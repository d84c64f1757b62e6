// MIND-SSC self-similarity descriptor fused with the block-mean pooling that always follows it.
//
// Replaces MINDSSC (reference src/convexAdam/convex_adam_utils.py:24-68) and the F.avg_pool3d(feat, grid_sp)
// calls at convex_adam_MIND.py:118-119,149-150.  The two one-hot dilated conv3d of the reference are pure
// shifts: only 6 distinct neighbours (+-dilation along each axis) are read per patch tap and the 12 channels
// are differences of pairs of them.  Two passes because of the global mean at utils.py:61 (no host sync here:
// the clamp bounds live in a device scalar):
//   phase A  per voxel: 12 patch-SSDs (sequential (2r+1)^3 sums, /k^3), minus channel-min, channel-mean -> ws
//   stats    deterministic two-level reduction of the channel-mean (double), clamp bounds lo/hi
//   phase B  per pooled voxel: exp(-mind/clamp(var)) summed over the gs^3 block in torch order, /gs^3
// Summation orders follow torch-CPU so that the only differences to the reference are expf (SLEEF vs CUDA
// libdevice, <= 2 ulp) and the global mean (double here, float cascade there).
#include "ca_platform.h"

struct MindTable {
    signed char ia[12], ib[12];   // output channel c = (nb[ia] - nb[ib])^2, nb = the 6-neighbourhood below
    signed char inv_perm[12];     // un-permuted channel o lives at output channel inv_perm[o]
    signed char off[6][3];        // neighbour offsets in units of the dilation, (H,W,D)
};

// utils.py:30-47: pairs (x>y) & (dist==2) in row-major order, then the permutation at :66
static MindTable make_mind_table() {
    static const int six[6][3] = {{0, 1, 1}, {1, 1, 0}, {1, 0, 1}, {1, 1, 2}, {2, 1, 1}, {1, 2, 1}};
    static const int perm[12] = {6, 8, 1, 11, 2, 10, 0, 7, 9, 4, 5, 3};
    MindTable t;
    int px[12], py[12], n = 0;
    for (int x = 0; x < 6; ++x)
        for (int y = 0; y < 6; ++y) {
            int dist = 0;
            for (int k = 0; k < 3; ++k) dist += (six[x][k] - six[y][k]) * (six[x][k] - six[y][k]);
            if (x > y && dist == 2) { px[n] = x; py[n] = y; ++n; }
        }
    for (int c = 0; c < 12; ++c) {
        t.ia[c] = (signed char)px[perm[c]];
        t.ib[c] = (signed char)py[perm[c]];
        t.inv_perm[perm[c]] = (signed char)c;
    }
    for (int k = 0; k < 6; ++k)
        for (int a = 0; a < 3; ++a) t.off[k][a] = (signed char)(six[k][a] - 1);
    return t;
}

__global__ void __launch_bounds__(256) k_mind_phase_a(const float* __restrict__ img, int H, int W, int D, int radius,
                                                      int dil, MindTable tb, float* __restrict__ mind,
                                                      float* __restrict__ var, double* __restrict__ partial) {
    __shared__ double red[256];
    const int64_t V = (int64_t)H * W * D;
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    if (o < V) {
        const int z = (int)(o % D), y = (int)((o / D) % W), x = (int)(o / ((int64_t)D * W));
        float acc[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) acc[c] = 0.f;
        for (int dx = -radius; dx <= radius; ++dx)
            for (int dy = -radius; dy <= radius; ++dy)
                for (int dz = -radius; dz <= radius; ++dz) {
                    const int tx = ca_clamp(x + dx, 0, H - 1), ty = ca_clamp(y + dy, 0, W - 1), tz = ca_clamp(z + dz, 0, D - 1);
                    float nb[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int qx = ca_clamp(tx + dil * tb.off[k][0], 0, H - 1), qy = ca_clamp(ty + dil * tb.off[k][1], 0, W - 1),
                                  qz = ca_clamp(tz + dil * tb.off[k][2], 0, D - 1);
                        nb[k] = __ldg(img + ((int64_t)qx * W + qy) * D + qz);
                    }
#pragma unroll
                    for (int c = 0; c < 12; ++c) {
                        const float df = __fsub_rn(nb[tb.ia[c]], nb[tb.ib[c]]);
                        acc[c] = __fadd_rn(acc[c], __fmul_rn(df, df));
                    }
                }
        const int k1 = 2 * radius + 1;
        const float k3 = (float)(k1 * k1 * k1);
        float mn = 0.f;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            acc[c] = ca_div(acc[c], k3);
            mn = (c == 0 || acc[c] < mn) ? acc[c] : mn;
        }
#pragma unroll
        for (int c = 0; c < 12; ++c) acc[c] = __fsub_rn(acc[c], mn);
        float s = 0.f;
#pragma unroll
        for (int oc = 0; oc < 12; ++oc) s = __fadd_rn(s, acc[tb.inv_perm[oc]]);   // torch.mean(mind, 1): un-permuted order
        v = ca_div(s, 12.f);
#pragma unroll
        for (int c = 0; c < 12; ++c) mind[c * V + o] = acc[c];
        var[o] = v;
    }
    red[threadIdx.x] = (o < V) ? (double)v : 0.0;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// stats[0] = lo, stats[1] = hi  (utils.py:61: clamp(mind_var, mean*0.001, mean*1000), mean via .item() -> double)
__global__ void __launch_bounds__(1024) k_mind_stats(const double* __restrict__ partial, int npartial, int64_t V,
                                                     float* __restrict__ stats) {
    __shared__ double red[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < npartial; i += 1024) s += partial[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 512; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float mean = (float)(red[0] / (double)V);
        stats[0] = (float)((double)mean * 0.001);
        stats[1] = (float)((double)mean * 1000.0);
    }
}

__global__ void __launch_bounds__(256) k_mind_phase_b(const float* __restrict__ mind, const float* __restrict__ var,
                                                      const float* __restrict__ stats, int H, int W, int D, int gs, int h,
                                                      int w, int d, float* __restrict__ out) {
    const int64_t V = (int64_t)H * W * D, n = (int64_t)h * w * d;
    const float lo = stats[0], hi = stats[1];
    const float cnt = (float)(gs * gs * gs);
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(o % d), y = (int)((o / d) % w), x = (int)(o / ((int64_t)d * w));
        float acc[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) acc[c] = 0.f;
        for (int a = 0; a < gs; ++a)
            for (int b = 0; b < gs; ++b)
                for (int e = 0; e < gs; ++e) {
                    const int64_t q = ((int64_t)(x * gs + a) * W + (y * gs + b)) * D + (z * gs + e);
                    float vv = __ldg(var + q);
                    vv = vv < lo ? lo : (vv > hi ? hi : vv);
#pragma unroll
                    for (int c = 0; c < 12; ++c) {
                        const float f = expf(-ca_div(__ldg(mind + c * V + q), vv));
                        acc[c] = __fadd_rn(acc[c], f);
                    }
                }
#pragma unroll
        for (int c = 0; c < 12; ++c) out[c * n + o] = (gs == 1) ? acc[c] : ca_div(acc[c], cnt);
    }
}

extern "C" size_t ca_mindssc_workspace_bytes(int H, int W, int D) {
    const int64_t V = (int64_t)H * W * D;
    return (size_t)(13 * V * 4 + 8 * ca_divup(V, 256) + 512);
}

extern "C" int ca_mindssc_pooled(const float* img, int H, int W, int D, int radius, int dilation, int gs_a, float* out_a,
                                 int gs_b, float* out_b, void* workspace, size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(img && workspace, "mindssc: null pointer");
    CA_CHECK_ARG(H > 0 && W > 0 && D > 0 && radius >= 0 && dilation >= 1, "mindssc: bad dims");
    CA_CHECK_ARG(gs_a >= 1 && out_a, "mindssc: first output is mandatory");
    CA_CHECK_ARG(gs_b == 0 || (gs_b >= 1 && out_b), "mindssc: bad second output");
    CA_CHECK_ARG(H / gs_a > 0 && W / gs_a > 0 && D / gs_a > 0, "mindssc: volume smaller than the pooling block");
    if (workspace_bytes < ca_mindssc_workspace_bytes(H, W, D)) {
        ca_set_error("mindssc: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t V = (int64_t)H * W * D;
    const int nblk = (int)ca_divup(V, 256);
    float* mind = (float*)workspace;
    float* var = mind + 12 * V;
    double* partial = (double*)(((uintptr_t)(var + V) + 255) & ~(uintptr_t)255);
    float* stats = (float*)(partial + nblk);
    const MindTable tb = make_mind_table();
    CA_LAUNCH(k_mind_phase_a, dim3(nblk), dim3(256), 0, st, img, H, W, D, radius, dilation, tb, mind, var, partial);
    CA_CHECK_LAUNCH("mind_phase_a");
    CA_LAUNCH(k_mind_stats, dim3(1), dim3(1024), 0, st, partial, nblk, V, stats);
    CA_CHECK_LAUNCH("mind_stats");
    for (int pass = 0; pass < 2; ++pass) {
        const int gs = pass ? gs_b : gs_a;
        float* out = pass ? out_b : out_a;
        if (gs == 0) continue;
        const int h = H / gs, w = W / gs, d = D / gs;
        if (h <= 0 || w <= 0 || d <= 0) { ca_set_error("mindssc: volume smaller than the pooling block"); return CA_ERR_BAD_ARGUMENT; }
        const int64_t n = (int64_t)h * w * d;
        int64_t nb = ca_divup(n, 256);
        if (nb > 148 * 32) nb = 148 * 32;
        CA_LAUNCH(k_mind_phase_b, dim3((unsigned)nb), dim3(256), 0, st, mind, var, stats, H, W, D, gs, h, w, d, out);
        CA_CHECK_LAUNCH("mind_phase_b");
    }
    return CA_OK;
}
