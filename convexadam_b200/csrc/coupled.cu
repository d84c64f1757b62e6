// Coupled-convex regularisation of the cost volume.
//
// Replaces `coupled_convex` (reference src/convexAdam/convex_adam_utils.py:93-109): six passes, each
//     sel(x)  = argmin_l ( ssd_l(x) + coeff * || mesh_l - soft(x) ||^2 )          (utils.py:104-105)
//     soft    = box3( mesh[:, sel] )                                                (utils.py:107)
// The reference walks 96 H-slices in a Python loop (~4000 tiny launches per direction); here one streaming kernel
// reads the whole (L,n) volume once per pass: HBM-bound (L*n*4 bytes), each thread owns 4 consecutive voxels
// (128-bit L1-bypassing loads, label stride n), the label range is split over gridDim.y so that the grid covers
// several waves of 148 SMs, and the partial minima meet in one 64-bit atomicMin per voxel on the key
// (cost bits << 32 | label): costs are >= +0 so their bit patterns order like the floats, and equal costs fall
// back to the lower label == torch.argmin's first-minimum rule.  The plain argmin of `correlate` (utils.py:87)
// is the same kernel without the regulariser.
// Arithmetic is the reference's, op by op: (m-s) rounded, squared rounded, ((q0+q1)+q2), *coeff, +ssd.
#include "ca_platform.h"

#include <cmath>

static __device__ __forceinline__ unsigned long long pack_key(float cost, int label) {
    return ((unsigned long long)__float_as_uint(cost) << 32) | (unsigned)label;
}

template <bool REG, int VEC>
__global__ void __launch_bounds__(256) k_cost_argmin(const float* __restrict__ ssd, int64_t n, int L,
                                                     const float* __restrict__ mesh, const float* __restrict__ soft,
                                                     float coeff, int labels_per_split, int use_atomic,
                                                     unsigned long long* __restrict__ keys) {
    CA_DYN_SMEM(float4, smesh);
    const int l0 = (int)blockIdx.y * labels_per_split;
    const int l1 = ca_min(L, l0 + labels_per_split);
    if (REG) {
        for (int i = threadIdx.x; i < l1 - l0; i += blockDim.x)
            smesh[i] = make_float4(mesh[l0 + i], mesh[L + l0 + i], mesh[2 * L + l0 + i], 0.f);
        __syncthreads();
    }
    const int64_t v0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (v0 >= n) return;
    float s0[VEC], s1[VEC], s2[VEC], best[VEC];
    int bl[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        best[i] = INFINITY;
        bl[i] = l0;
        if (REG) { s0[i] = soft[v0 + i]; s1[i] = soft[n + v0 + i]; s2[i] = soft[2 * n + v0 + i]; }
    }
    const float* sp = ssd + v0;
#pragma unroll 4
    for (int l = l0; l < l1; ++l) {
        float c[VEC];
        if (VEC == 4) {
            const float4 c4 = ca_ld_stream4(reinterpret_cast<const float4*>(sp + (int64_t)l * n));
            c[0] = c4.x; c[1 % VEC] = c4.y; c[2 % VEC] = c4.z; c[3 % VEC] = c4.w;
        } else {
            c[0] = ca_ld_stream1(sp + (int64_t)l * n);
        }
        float4 mv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (REG) mv = smesh[l - l0];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float cost = c[i];
            if (REG) {
                const float d0 = __fsub_rn(mv.x, s0[i]), d1 = __fsub_rn(mv.y, s1[i]), d2 = __fsub_rn(mv.z, s2[i]);
                const float dist = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
                cost = __fadd_rn(c[i], __fmul_rn(coeff, dist));
            }
            if (cost < best[i]) { best[i] = cost; bl[i] = l; }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const unsigned long long k = pack_key(best[i], bl[i]);
        if (use_atomic) atomicMin(keys + v0 + i, k);
        else keys[v0 + i] = k;
    }
}

// soft[c](x) = box3( mesh[c, sel] )(x): ascending (dH,dW,dD) sum of the in-volume taps, / 27  (utils.py:96,107)
__global__ void __launch_bounds__(256) k_soft_from_keys(const unsigned long long* __restrict__ keys,
                                                        const float* __restrict__ mesh, int h, int w, int d, int L,
                                                        float* __restrict__ soft) {
    const int64_t n = (int64_t)h * w * d;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(o % d), y = (int)((o / d) % w), x = (int)(o / ((int64_t)d * w));
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int xx = ca_max(x - 1, 0); xx < ca_min(x + 2, h); ++xx)
            for (int yy = ca_max(y - 1, 0); yy < ca_min(y + 2, w); ++yy)
                for (int zz = ca_max(z - 1, 0); zz < ca_min(z + 2, d); ++zz) {
                    const int l = (int)(unsigned)(keys[((int64_t)xx * w + yy) * d + zz] & 0xffffffffull);
                    a0 = __fadd_rn(a0, __ldg(mesh + l));
                    a1 = __fadd_rn(a1, __ldg(mesh + L + l));
                    a2 = __fadd_rn(a2, __ldg(mesh + 2 * L + l));
                }
        soft[o] = ca_div(a0, 27.f);
        soft[n + o] = ca_div(a1, 27.f);
        soft[2 * n + o] = ca_div(a2, 27.f);
    }
}

__global__ void __launch_bounds__(256) k_keys_to_i64(const unsigned long long* __restrict__ keys, int64_t n,
                                                     int64_t* __restrict__ out) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x)
        out[o] = (int64_t)(keys[o] & 0xffffffffull);
}
__global__ void __launch_bounds__(256) k_i64_to_keys(const int64_t* __restrict__ in, int64_t n,
                                                     unsigned long long* __restrict__ keys) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x)
        keys[o] = (unsigned long long)in[o] & 0xffffffffull;
}

static inline unsigned small_grid(int64_t n) {
    int64_t b = ca_divup(n, 256);
    if (b > 148 * 16) b = 148 * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

// Enqueues one full pass over the cost volume.  mesh == nullptr: plain argmin (no regulariser).
int ca_launch_cost_argmin(const float* ssd, int64_t n, int L, const float* mesh, const float* soft, float coeff,
                          unsigned long long* keys, cudaStream_t st) {
    const bool reg = mesh != nullptr;
    const int vec = (n % 4 == 0 && (((uintptr_t)ssd) & 15) == 0) ? 4 : 1;
    const int64_t nthreads = ca_divup(n, vec);
    const unsigned gx = (unsigned)ca_divup(nthreads, 256);
    // aim for >= 4 waves of 148 SMs x 8 resident CTAs; the regulariser table must also fit 48 KB of shared memory
    int nsplit = (int)ca_divup(4 * 148 * 8, gx);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > 16) nsplit = 16;
    int lps = (int)ca_divup(L, nsplit);
    if (lps > 3072) lps = 3072;
    nsplit = (int)ca_divup(L, lps);
    const int use_atomic = nsplit > 1;
    if (use_atomic) {
        if (cudaMemsetAsync(keys, 0xff, (size_t)n * 8, st) != cudaSuccess) { ca_set_error("cost_argmin: memset failed"); return CA_ERR_LAUNCH; }
    }
    const size_t smem = reg ? (size_t)lps * sizeof(float4) : 0;
    dim3 grid(gx, (unsigned)nsplit), block(256);
    if (reg && vec == 4) { auto k = k_cost_argmin<true, 4>; CA_LAUNCH(k, grid, block, smem, st, ssd, n, L, mesh, soft, coeff, lps, use_atomic, keys); }
    else if (reg) { auto k = k_cost_argmin<true, 1>; CA_LAUNCH(k, grid, block, smem, st, ssd, n, L, mesh, soft, coeff, lps, use_atomic, keys); }
    else if (vec == 4) { auto k = k_cost_argmin<false, 4>; CA_LAUNCH(k, grid, block, smem, st, ssd, n, L, mesh, soft, coeff, lps, use_atomic, keys); }
    else { auto k = k_cost_argmin<false, 1>; CA_LAUNCH(k, grid, block, smem, st, ssd, n, L, mesh, soft, coeff, lps, use_atomic, keys); }
    CA_CHECK_LAUNCH("cost_argmin");
    return CA_OK;
}

int ca_launch_keys_to_i64(const unsigned long long* keys, int64_t n, int64_t* out, cudaStream_t st) {
    CA_LAUNCH(k_keys_to_i64, dim3(small_grid(n)), dim3(256), 0, st, keys, n, out);
    CA_CHECK_LAUNCH("keys_to_i64");
    return CA_OK;
}

extern "C" size_t ca_coupled_convex_workspace_bytes(int h, int w, int d, int disp_hw) {
    (void)disp_hw;
    const size_t n = (size_t)h * w * d;
    return n * 8 + 256;
}

extern "C" int ca_coupled_convex(const float* ssd, const int64_t* argmin, const float* mesh, int h, int w, int d,
                                 int disp_hw, float* disp_soft, int64_t* sel_out, void* workspace, size_t workspace_bytes,
                                 ca_stream_t stream) {
    CA_CHECK_ARG(ssd && argmin && mesh && disp_soft && workspace, "coupled_convex: null pointer");
    CA_CHECK_ARG(h > 0 && w > 0 && d > 0 && disp_hw >= 1, "coupled_convex: bad dims");
    if (workspace_bytes < ca_coupled_convex_workspace_bytes(h, w, d, disp_hw)) {
        ca_set_error("coupled_convex: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    static const float coeffs[6] = {0.003f, 0.01f, 0.03f, 0.1f, 0.3f, 1.f};   // utils.py:98
    cudaStream_t st = (cudaStream_t)stream;
    const int S = 2 * disp_hw + 1, L = S * S * S;
    const int64_t n = (int64_t)h * w * d;
    unsigned long long* keys = (unsigned long long*)workspace;
    CA_LAUNCH(k_i64_to_keys, dim3(small_grid(n)), dim3(256), 0, st, argmin, n, keys);
    CA_LAUNCH(k_soft_from_keys, dim3(small_grid(n)), dim3(256), 0, st, keys, mesh, h, w, d, L, disp_soft);
    CA_CHECK_LAUNCH("coupled_convex/init");
    for (int j = 0; j < 6; ++j) {
        int rc = ca_launch_cost_argmin(ssd, n, L, mesh, disp_soft, coeffs[j], keys, st);
        if (rc != CA_OK) return rc;
        if (sel_out) {
            rc = ca_launch_keys_to_i64(keys, n, sel_out + (int64_t)j * n, st);
            if (rc != CA_OK) return rc;
        }
        CA_LAUNCH(k_soft_from_keys, dim3(small_grid(n)), dim3(256), 0, st, keys, mesh, h, w, d, L, disp_soft);
        CA_CHECK_LAUNCH("coupled_convex/soft");
    }
    return CA_OK;
}

// ---- the same stage in pieces, for a pair that is split into H slabs over several GPUs --------------------------------
// The host runs the loop of utils.py:98-107 itself and exchanges one plane of selected labels with the neighbouring
// slabs before each smoothing step (convexadam_b200/slab.py).

// soft[c](x) = box3(mesh[c, sel])(x) for the planes [1, P-1) of a (P, w, d) window of labels; a negative label marks a
// plane outside the volume (zero padding of avg_pool3d: the tap is skipped).
__global__ void __launch_bounds__(256) k_soft_from_label_window(const int64_t* __restrict__ sel, const float* __restrict__ mesh,
                                                                int P, int w, int d, int L, float* __restrict__ soft) {
    const int64_t n = (int64_t)(P - 2) * w * d;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(o % d), y = (int)((o / d) % w), x = (int)(o / ((int64_t)d * w)) + 1;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int xx = x - 1; xx < x + 2; ++xx)
            for (int yy = ca_max(y - 1, 0); yy < ca_min(y + 2, w); ++yy)
                for (int zz = ca_max(z - 1, 0); zz < ca_min(z + 2, d); ++zz) {
                    const int64_t l = sel[((int64_t)xx * w + yy) * d + zz];
                    if (l < 0) continue;
                    a0 = __fadd_rn(a0, __ldg(mesh + l));
                    a1 = __fadd_rn(a1, __ldg(mesh + L + l));
                    a2 = __fadd_rn(a2, __ldg(mesh + 2 * L + l));
                }
        soft[o] = ca_div(a0, 27.f);
        soft[n + o] = ca_div(a1, 27.f);
        soft[2 * n + o] = ca_div(a2, 27.f);
    }
}

extern "C" size_t ca_cc_select_workspace_bytes(int64_t n) { return (size_t)n * 8 + 256; }

extern "C" int ca_cc_select(const float* ssd, int64_t n, int disp_hw, const float* mesh, const float* soft, float coeff,
                            int64_t* sel, void* workspace, size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(ssd && sel && workspace && n > 0 && disp_hw >= 1, "cc_select: bad argument");
    CA_CHECK_ARG((mesh == nullptr) == (soft == nullptr), "cc_select: mesh and soft go together");
    if (workspace_bytes < ca_cc_select_workspace_bytes(n)) {
        ca_set_error("cc_select: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    const int S = 2 * disp_hw + 1;
    unsigned long long* keys = (unsigned long long*)workspace;
    int rc = ca_launch_cost_argmin(ssd, n, S * S * S, mesh, soft, coeff, keys, (cudaStream_t)stream);
    if (rc != CA_OK) return rc;
    return ca_launch_keys_to_i64(keys, n, sel, (cudaStream_t)stream);
}

extern "C" int ca_cc_smooth(const int64_t* sel_window, const float* mesh, int P, int w, int d, int disp_hw, float* soft,
                            ca_stream_t stream) {
    CA_CHECK_ARG(sel_window && mesh && soft && P >= 3 && w > 0 && d > 0 && disp_hw >= 1, "cc_smooth: bad argument");
    const int S = 2 * disp_hw + 1;
    const int64_t n = (int64_t)(P - 2) * w * d;
    CA_LAUNCH(k_soft_from_label_window, dim3(small_grid(n)), dim3(256), 0, (cudaStream_t)stream, sel_window, mesh, P, w, d, S * S * S, soft);
    CA_CHECK_LAUNCH("cc_smooth");
    return CA_OK;
}
