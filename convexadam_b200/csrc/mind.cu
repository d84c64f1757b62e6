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

#include <cstdlib>

// The 6-neighbourhood (units of the dilation, (H,W,D)) and, per OUTPUT channel c (after the permutation at utils.py:66),
// the two neighbours whose difference is squared: utils.py:30-47 selects the pairs (x>y) & (dist==2) in row-major
// order.  Hard-wired so that every index is a compile-time constant (registers, no local memory);
// mind_table_is_consistent() re-derives them the reference's way and ca_mindssc_pooled refuses to run on a mismatch.
#define MIND_IA {4, 5, 2, 5, 2, 5, 1, 4, 5, 3, 4, 3}
#define MIND_IB {2, 0, 0, 4, 1, 3, 0, 3, 1, 2, 1, 0}
#define MIND_INV {6, 2, 4, 11, 9, 10, 0, 7, 1, 8, 5, 3}   /* un-permuted channel o lives at output channel INV[o] */
#define MIND_OFF {{-1, 0, 0}, {0, 0, -1}, {0, -1, 0}, {0, 0, 1}, {1, 0, 0}, {0, 1, 0}}

static bool mind_table_is_consistent() {
    static const int six[6][3] = {{0, 1, 1}, {1, 1, 0}, {1, 0, 1}, {1, 1, 2}, {2, 1, 1}, {1, 2, 1}};
    static const int perm[12] = {6, 8, 1, 11, 2, 10, 0, 7, 9, 4, 5, 3};
    const int ia[12] = MIND_IA, ib[12] = MIND_IB, inv[12] = MIND_INV, off[6][3] = MIND_OFF;
    int px[12], py[12], n = 0;
    for (int x = 0; x < 6; ++x)
        for (int y = 0; y < 6; ++y) {
            int dist = 0;
            for (int k = 0; k < 3; ++k) dist += (six[x][k] - six[y][k]) * (six[x][k] - six[y][k]);
            if (x > y && dist == 2 && n < 12) { px[n] = x; py[n] = y; ++n; }
        }
    if (n != 12) return false;
    for (int c = 0; c < 12; ++c)
        if (ia[c] != px[perm[c]] || ib[c] != py[perm[c]] || inv[perm[c]] != c) return false;
    for (int k = 0; k < 6; ++k)
        for (int a = 0; a < 3; ++a)
            if (off[k][a] != six[k][a] - 1) return false;
    return true;
}

// grid (ceil(D/128), W, H), block 128 along D: coalesced rows, no 64-bit index maths.
// R > 0: the patch radius is a compile-time constant and the replicate-padded (clamped) coordinates of every tap and of
// its +-dilation neighbours are formed once per voxel and axis -- the kernel is bound by issue slots, and per-load clamping
// was most of its instructions (ncu: 3350 instructions per voxel, 56 % ALU pipe).  R == 0: any radius, clamps per load.
template <int R>
__global__ void __launch_bounds__(128) k_mind_phase_a(const float* __restrict__ img, int H, int W, int D, int radius,
                                                      int dil, float* __restrict__ mind, float* __restrict__ var,
                                                      double* __restrict__ partial) {
    __shared__ double red[128];
    constexpr int IA[12] = MIND_IA, IB[12] = MIND_IB, INV[12] = MIND_INV, OFF[6][3] = MIND_OFF;
    const int64_t V = (int64_t)H * W * D;
    const int z = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, x = (int)blockIdx.z;
    float v = 0.f;
    if (z < D) {
        float acc[12];
#pragma unroll
        for (int c = 0; c < 12; ++c) acc[c] = 0.f;
        if (R > 0) {
            constexpr int K = 2 * R + 1;
            // element offsets of clamp(clamp(x + t) + o * dil) for t in [-R, R], o in {-1, 0, +1}, per axis
            int xo[K][3], yo[K][3], zo[K][3];
#pragma unroll
            for (int t = 0; t < K; ++t)
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    xo[t][o] = ca_clamp(ca_clamp(x + t - R, 0, H - 1) + (o - 1) * dil, 0, H - 1) * W * D;
                    yo[t][o] = ca_clamp(ca_clamp(y + t - R, 0, W - 1) + (o - 1) * dil, 0, W - 1) * D;
                    zo[t][o] = ca_clamp(ca_clamp(z + t - R, 0, D - 1) + (o - 1) * dil, 0, D - 1);
                }
#pragma unroll
            for (int dx = 0; dx < K; ++dx)
#pragma unroll
                for (int dy = 0; dy < K; ++dy)
#pragma unroll
                    for (int dz = 0; dz < K; ++dz) {
                        float nb[6];
#pragma unroll
                        for (int k = 0; k < 6; ++k) nb[k] = __ldg(img + xo[dx][OFF[k][0] + 1] + yo[dy][OFF[k][1] + 1] + zo[dz][OFF[k][2] + 1]);
#pragma unroll
                        for (int c = 0; c < 12; ++c) {
                            const float df = __fsub_rn(nb[IA[c]], nb[IB[c]]);
                            acc[c] = __fadd_rn(acc[c], __fmul_rn(df, df));
                        }
                    }
        } else {
            for (int dx = -radius; dx <= radius; ++dx)
                for (int dy = -radius; dy <= radius; ++dy)
                    for (int dz = -radius; dz <= radius; ++dz) {
                        const int tx = ca_clamp(x + dx, 0, H - 1), ty = ca_clamp(y + dy, 0, W - 1), tz = ca_clamp(z + dz, 0, D - 1);
                        float nb[6];
#pragma unroll
                        for (int k = 0; k < 6; ++k) {
                            const int qx = ca_clamp(tx + dil * OFF[k][0], 0, H - 1), qy = ca_clamp(ty + dil * OFF[k][1], 0, W - 1),
                                      qz = ca_clamp(tz + dil * OFF[k][2], 0, D - 1);
                            nb[k] = __ldg(img + ((int64_t)qx * W + qy) * D + qz);
                        }
#pragma unroll
                        for (int c = 0; c < 12; ++c) {
                            const float df = __fsub_rn(nb[IA[c]], nb[IB[c]]);
                            acc[c] = __fadd_rn(acc[c], __fmul_rn(df, df));
                        }
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
        for (int oc = 0; oc < 12; ++oc) s = __fadd_rn(s, acc[INV[oc]]);   // torch.mean(mind, 1): un-permuted order
        v = ca_div(s, 12.f);
        const int64_t o = ((int64_t)x * W + y) * D + z;
#pragma unroll
        for (int c = 0; c < 12; ++c) mind[c * V + o] = acc[c];
        var[o] = v;
    }
    red[threadIdx.x] = (z < D) ? (double)v : 0.0;
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = red[0];
}

// The same phase for NR consecutive rows per thread (grid (ceil(D/128), ceil(W/NR), H)): the squared differences at a tap
// location are the same for every voxel whose patch contains it, so a thread that owns the voxels (x, y0 .. y0+NR-1, z) forms
// them once per location row and adds them to the (up to 2R+1) voxels of its column they belong to -- 6 loads and 24 flops per
// location instead of per (voxel, tap): 2 400 -> 810 instructions per voxel at R = 1, NR = 4.  Every accumulator still receives
// its (2R+1)^3 terms in torch's order (dx, dy, dz ascending: the location rows are walked upwards, which is dy ascending for
// each voxel), and the block sums of the channel mean are formed per row exactly as k_mind_phase_a forms them, so mind, var and
// the clamp bounds are bit-identical to that kernel's.
template <int R, int NR, int MINB>
__global__ void __launch_bounds__(128, MINB) k_mind_phase_a_rows(const float* __restrict__ img, int H, int W, int D, int dil,
                                                           float* __restrict__ mind, float* __restrict__ var,
                                                           double* __restrict__ partial) {
    __shared__ double red[128];
    constexpr int IA[12] = MIND_IA, IB[12] = MIND_IB, INV[12] = MIND_INV, OFF[6][3] = MIND_OFF;
    constexpr int K = 2 * R + 1, NL = NR + 2 * R;
    const int64_t V = (int64_t)H * W * D;
    const int z = (int)(blockIdx.x * blockDim.x + threadIdx.x), y0 = (int)blockIdx.y * NR, x = (int)blockIdx.z;
    float vrow[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) vrow[i] = 0.f;
    if (z < D) {
        float acc[NR][12];
#pragma unroll
        for (int i = 0; i < NR; ++i)
#pragma unroll
            for (int c = 0; c < 12; ++c) acc[i][c] = 0.f;
        int xo[K][3], zo[K][3], yo[NL][3];
#pragma unroll
        for (int t = 0; t < K; ++t)
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                xo[t][o] = ca_clamp(ca_clamp(x + t - R, 0, H - 1) + (o - 1) * dil, 0, H - 1) * W * D;
                zo[t][o] = ca_clamp(ca_clamp(z + t - R, 0, D - 1) + (o - 1) * dil, 0, D - 1);
            }
#pragma unroll
        for (int j = 0; j < NL; ++j)
#pragma unroll
            for (int o = 0; o < 3; ++o) yo[j][o] = ca_clamp(ca_clamp(y0 + j - R, 0, W - 1) + (o - 1) * dil, 0, W - 1) * D;
#pragma unroll
        for (int dx = 0; dx < K; ++dx)
#pragma unroll
            for (int j = 0; j < NL; ++j)          // location row y0 + j - R: tap dy = j - i of the voxel in row y0 + i
#pragma unroll
                for (int dz = 0; dz < K; ++dz) {
                    float nb[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) nb[k] = __ldg(img + xo[dx][OFF[k][0] + 1] + yo[j][OFF[k][1] + 1] + zo[dz][OFF[k][2] + 1]);
#pragma unroll
                    for (int c = 0; c < 12; ++c) {
                        const float df = __fsub_rn(nb[IA[c]], nb[IB[c]]);
                        const float d2 = __fmul_rn(df, df);
#pragma unroll
                        for (int i = 0; i < NR; ++i)
                            if (j - i >= 0 && j - i < K) acc[i][c] = __fadd_rn(acc[i][c], d2);
                    }
                }
        const float k3 = (float)(K * K * K);
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            if (y0 + i >= W) break;
            float mn = 0.f;
#pragma unroll
            for (int c = 0; c < 12; ++c) {
                acc[i][c] = ca_div(acc[i][c], k3);
                mn = (c == 0 || acc[i][c] < mn) ? acc[i][c] : mn;
            }
#pragma unroll
            for (int c = 0; c < 12; ++c) acc[i][c] = __fsub_rn(acc[i][c], mn);
            float s = 0.f;
#pragma unroll
            for (int oc = 0; oc < 12; ++oc) s = __fadd_rn(s, acc[i][INV[oc]]);   // torch.mean(mind, 1): un-permuted order
            vrow[i] = ca_div(s, 12.f);
            const int64_t o = ((int64_t)x * W + (y0 + i)) * D + z;
#pragma unroll
            for (int c = 0; c < 12; ++c) mind[c * V + o] = acc[i][c];
            var[o] = vrow[i];
        }
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) {
        if (y0 + i >= W) break;                                     // uniform over the CTA
        red[threadIdx.x] = (z < D) ? (double)vrow[i] : 0.0;
        __syncthreads();
        for (int s = 64; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) partial[((int64_t)x * W + (y0 + i)) * gridDim.x + blockIdx.x] = red[0];
        __syncthreads();
    }
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

// grid (ceil(d/128), w, h), block 128 along the pooled D axis
__global__ void __launch_bounds__(128) k_mind_phase_b(const float* __restrict__ mind, const float* __restrict__ var,
                                                      const float* __restrict__ stats, int H, int W, int D, int gs, int h,
                                                      int w, int d, float* __restrict__ out) {
    const int64_t V = (int64_t)H * W * D, n = (int64_t)h * w * d;
    const float lo = stats[0], hi = stats[1];
    const float cnt = (float)(gs * gs * gs);
    const int z = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, x = (int)blockIdx.z;
    if (z >= d) return;
    float acc[12];
#pragma unroll
    for (int c = 0; c < 12; ++c) acc[c] = 0.f;
    if (gs == 2 && (D & 1) == 0) {
        // the common pooling factor: the two voxels along D come as one aligned 64-bit load (coalesced across the warp),
        // same summation order (a, b, e ascending) as the generic loop below
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int64_t q = ((int64_t)(x * 2 + a) * W + (y * 2 + b)) * D + z * 2;
                float2 vv = __ldg(reinterpret_cast<const float2*>(var + q));
                vv.x = vv.x < lo ? lo : (vv.x > hi ? hi : vv.x);
                vv.y = vv.y < lo ? lo : (vv.y > hi ? hi : vv.y);
                float2 mm[12];
#pragma unroll
                for (int c = 0; c < 12; ++c) mm[c] = __ldg(reinterpret_cast<const float2*>(mind + c * V + q));
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    acc[c] = __fadd_rn(acc[c], expf(-ca_div(mm[c].x, vv.x)));
                    acc[c] = __fadd_rn(acc[c], expf(-ca_div(mm[c].y, vv.y)));
                }
            }
    } else {
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
    }
    const int64_t o = ((int64_t)x * w + y) * d + z;
#pragma unroll
    for (int c = 0; c < 12; ++c) out[c * n + o] = (gs == 1) ? acc[c] : ca_div(acc[c], cnt);
}

extern "C" size_t ca_mindssc_workspace_bytes(int H, int W, int D) {
    const int64_t V = (int64_t)H * W * D;
    return (size_t)(13 * V * 4 + 8 * ((int64_t)H * W * ca_divup(D, 128)) + 1024);
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
    CA_CHECK_ARG(W < 65536 && H < 65536, "mindssc: volume too large");
    if (!mind_table_is_consistent()) { ca_set_error("mindssc: internal pair table inconsistent"); return CA_ERR_UNSUPPORTED; }
    const int nzb = (int)ca_divup(D, 128);
    const int64_t nblk64 = (int64_t)H * W * nzb;
    CA_CHECK_ARG(nblk64 < (1ll << 31), "mindssc: volume too large");
    const int nblk = (int)nblk64;
    float* mind = (float*)workspace;
    float* var = mind + 12 * V;
    double* partial = (double*)(((uintptr_t)(var + V) + 255) & ~(uintptr_t)255);
    float* stats = (float*)(partial + nblk);
    // element offsets are 32-bit in the specialised kernels
    const bool small = V < (1ll << 31);
    const dim3 agrid((unsigned)nzb, (unsigned)W, (unsigned)H);
    // CA_MIND_ROWS=1: one voxel per thread (k_mind_phase_a); default: 4 rows per thread.  Read per call (the parity test compares both).
    const char* rows_e = getenv("CA_MIND_ROWS");
    const int rows_env = rows_e && *rows_e ? atoi(rows_e) : 4;
#define CA_MIND_ROWS_LAUNCH(RR, NRR, MB) { auto k = k_mind_phase_a_rows<RR, NRR, MB>; const dim3 rgrid((unsigned)nzb, (unsigned)ca_divup(W, NRR), (unsigned)H); \
        CA_LAUNCH(k, rgrid, dim3(128), 0, st, img, H, W, D, dilation, mind, var, partial); }
    // measured per image at config 2 (phase A + stats + phase B): one voxel per thread 0.833 ms; 4 rows per thread compiled for
    // 4 / 5 / 6 resident CTAs 0.80 / 0.72 / 0.73 ms (unconstrained: 106 registers, 0.78 ms); 2 / 3 / 8 rows 0.775 / 0.80 / 1.20 ms
    if (radius == 1 && small && rows_env == 4) CA_MIND_ROWS_LAUNCH(1, 4, 5)
    else if (radius == 2 && small && rows_env == 4) CA_MIND_ROWS_LAUNCH(2, 4, 2)
    else if (radius == 1 && small) { auto k = k_mind_phase_a<1>; CA_LAUNCH(k, agrid, dim3(128), 0, st, img, H, W, D, radius, dilation, mind, var, partial); }
    else if (radius == 2 && small) { auto k = k_mind_phase_a<2>; CA_LAUNCH(k, agrid, dim3(128), 0, st, img, H, W, D, radius, dilation, mind, var, partial); }
    else { auto k = k_mind_phase_a<0>; CA_LAUNCH(k, agrid, dim3(128), 0, st, img, H, W, D, radius, dilation, mind, var, partial); }
    CA_CHECK_LAUNCH("mind_phase_a");
    CA_LAUNCH(k_mind_stats, dim3(1), dim3(1024), 0, st, partial, nblk, V, stats);
    CA_CHECK_LAUNCH("mind_stats");
    for (int pass = 0; pass < 2; ++pass) {
        const int gs = pass ? gs_b : gs_a;
        float* out = pass ? out_b : out_a;
        if (gs == 0) continue;
        const int h = H / gs, w = W / gs, d = D / gs;
        if (h <= 0 || w <= 0 || d <= 0) { ca_set_error("mindssc: volume smaller than the pooling block"); return CA_ERR_BAD_ARGUMENT; }
        CA_LAUNCH(k_mind_phase_b, dim3((unsigned)ca_divup(d, 128), (unsigned)w, (unsigned)h), dim3(128), 0, st, mind, var, stats, H, W, D,
                  gs, h, w, d, out);
        CA_CHECK_LAUNCH("mind_phase_b");
    }
    return CA_OK;
}
