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
#include <cstdlib>

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
// grid = (ceil(d/128), w, h): no 64-bit index maths; the label is the low word of the key.
__global__ void __launch_bounds__(128) k_soft_from_keys(const unsigned long long* __restrict__ keys,
                                                        const float* __restrict__ mesh, int h, int w, int d, int L,
                                                        float* __restrict__ soft) {
    const int z = (int)(blockIdx.x * blockDim.x + threadIdx.x), y = (int)blockIdx.y, x = (int)blockIdx.z;
    if (z >= d) return;
    const int64_t n = (int64_t)h * w * d;
    const unsigned* lab = reinterpret_cast<const unsigned*>(keys);       // little endian: label = low 32 bits of the key
    // all 27 label loads go out together (clamped coordinates), then the 81 mesh values; a tap outside the volume adds +0,
    // which leaves the ascending (dH,dW,dD) sum of the in-volume taps unchanged bit for bit
    int lbl[27];
    bool in[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
        const int xx = x + t / 9 - 1, yy = y + (t / 3) % 3 - 1, zz = z + t % 3 - 1;
        in[t] = xx >= 0 && xx < h && yy >= 0 && yy < w && zz >= 0 && zz < d;
        const int64_t q = ((int64_t)ca_min(ca_max(xx, 0), h - 1) * w + ca_min(ca_max(yy, 0), w - 1)) * d + ca_min(ca_max(zz, 0), d - 1);
        lbl[t] = (int)__ldg(lab + 2 * q);
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
        const float m0 = __ldg(mesh + lbl[t]), m1 = __ldg(mesh + L + lbl[t]), m2 = __ldg(mesh + 2 * L + lbl[t]);
        a0 = __fadd_rn(a0, in[t] ? m0 : 0.f);
        a1 = __fadd_rn(a1, in[t] ? m1 : 0.f);
        a2 = __fadd_rn(a2, in[t] ? m2 : 0.f);
    }
    const int64_t o = ((int64_t)x * w + y) * d + z;
    soft[o] = ca_div(a0, 27.f);
    soft[n + o] = ca_div(a1, 27.f);
    soft[2 * n + o] = ca_div(a2, 27.f);
}

__global__ void __launch_bounds__(256) k_keys_to_i64(const unsigned long long* __restrict__ keys, int64_t n,
                                                     int64_t* __restrict__ out) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x)
        out[o] = (int64_t)(keys[o] & 0xffffffffull);
}
// caller-supplied labels (the public coupled_convex operator takes any int64 tensor): clamped into [0, L) so that the
// pruned passes, which read ssd[label * n + v], can never leave the volume (the reference would raise an IndexError)
__global__ void __launch_bounds__(256) k_i64_to_keys(const int64_t* __restrict__ in, int64_t n, int L,
                                                     unsigned long long* __restrict__ keys) {
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = in[o];
        keys[o] = (unsigned long long)(l < 0 ? 0 : (l >= L ? L - 1 : l));
    }
}

// ---- certified selection from an approximate cost volume -------------------------------------------------------------
// ca_correlate_fast produces ssd' with |ssd' - ssd| <= 98 u * ssd (u = 2^-24): both are sums of the same non-negative
// terms, the reference's ordered arithmetic rounds at most 68 times per entry (14 for the channel sum, 27 per box
// filter), the fast one at most 30.  Adding the regulariser (computed identically) and the final rounding, every
// coupled cost satisfies |c' - c| <= eps * max(c, c') with eps = 100 u = 5.96e-6.  If l* is the reference's selection
// and l' the approximate one, c'(l*) <= c'(l') (1+eps)/(1-eps) = c'(l') (1 + 1.2e-5) < c'(l') * (1 + 2^-16): the reference's
// selection always lies in the band.  A voxel with a single label in the band is therefore decided; the others (a few
// hundred per million) are re-decided by k_cc_fixup, which evaluates the candidates' costs in the reference's exact
// operation order straight from the feature volumes.  The selections -- and with them disp_soft -- are the
// reference's bit for bit; only the cost volume itself is approximate.
// (The relative bound assumes no term df^2 underflows to a subnormal or to zero -- underflow would add an absolute error
// the band does not cover.  MIND descriptors cannot get there: every voxel has a channel equal to exp(0) = 1, differences
// below 2^-63 do not occur between values in [0,1] quantised to 2^-24, and an exactly zero cost is exact in both
// arithmetics.  Weighted one-hot label features (values 0 or ~10) cannot either; that path takes the exact volume anyway.)
static __device__ __forceinline__ float ca_band_hi(float c) { return __fmul_ru(c, 1.0f + 0x1p-16f); }

struct FixupArgs {
    const float* ssd;           // (L, n) approximate cost volume
    const float* mesh;          // (3, L)
    const float* soft;          // (3, n)
    const float* fix;           // (C, h, w, d) features the volume was built from
    const float* mov;
    const int* flagged;
    const int* nflagged;
    unsigned long long* keys;
    int64_t n;
    float coeff;
    int C, h, w, d, hw;
    int x0;                     // first H plane of the slab the (L, n) arrays describe (0: the whole volume)
};

// One CTA per flagged voxel: (1) the minimum approximate cost over ALL labels, (2) the labels within the band of it,
// (3) one warp per candidate evaluates the exact cost: raw SSD on the 5x5x5 neighbourhood (utils.py:84: (f-m) rounded,
// squared rounded, channels summed sequentially; the moving features zero-padded, utils.py:76), the first box filter at
// the 27 neighbours (ascending (dH,dW,dD) sum of the in-volume taps, /27; zero outside the volume), the second one at
// the voxel (utils.py:85), then + coeff * |mesh - soft|^2 as in utils.py:104.  The minimum of the (cost, label) keys is
// the first minimum in label order, whatever order the candidates were evaluated in.
static __device__ __forceinline__ float cc_cost(const FixupArgs& a, int L, int l, int v, float s0, float s1, float s2) {
    const float d0 = __fsub_rn(__ldg(a.mesh + l), s0), d1 = __fsub_rn(__ldg(a.mesh + L + l), s1), d2 = __fsub_rn(__ldg(a.mesh + 2 * L + l), s2);
    const float dist = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
    return __fadd_rn(ca_ld_stream1(a.ssd + (int64_t)l * a.n + v), __fmul_rn(a.coeff, dist));
}

#define CA_FIXUP_MAXL 4913   // (2*8+1)^3
__global__ void __launch_bounds__(128) k_cc_fixup(FixupArgs a) {
    __shared__ float s_raw[4][128];
    __shared__ float s_box[4][32];
    __shared__ float s_min[4];
    __shared__ unsigned long long s_key[4];
    __shared__ int s_ncand;
    __shared__ int s_cand[CA_FIXUP_MAXL];
    const int tid = (int)threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int S = 2 * a.hw + 1, L = S * S * S;
    const int count = *a.nflagged;
    float* raw = s_raw[wib];
    float* box = s_box[wib];
    const int64_t cs = (int64_t)a.h * a.w * a.d;
    for (int idx = (int)blockIdx.x; idx < count; idx += (int)gridDim.x) {
        const int v = a.flagged[idx];
        const int z = v % a.d, y = (v / a.d) % a.w, x = v / (a.d * a.w) + a.x0;
        const float s0 = a.soft[v], s1 = a.soft[a.n + v], s2 = a.soft[2 * a.n + v];
        if (tid == 0) s_ncand = 0;
        float mn = INFINITY;
#pragma unroll 6
        for (int l = tid; l < L; l += 128) mn = fminf(mn, cc_cost(a, L, l, v, s0, s1, s2));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (lane == 0) s_min[wib] = mn;
        __syncthreads();
        const float thr = ca_band_hi(fminf(fminf(s_min[0], s_min[1]), fminf(s_min[2], s_min[3])));
#pragma unroll 6
        for (int l = tid; l < L; l += 128)
            if (cc_cost(a, L, l, v, s0, s1, s2) <= thr) s_cand[atomicAdd(&s_ncand, 1)] = l;
        __syncthreads();
        const int ncand = s_ncand;
        unsigned long long best = ~0ull;
        for (int ci = wib; ci < ncand; ci += 4) {
            const int lc = s_cand[ci];
            const int kH = lc % S - a.hw, kW = (lc / S) % S - a.hw, kD = lc / (S * S) - a.hw;
            // raw SSD at x + [-2,2]^3 (0 outside the volume: a +0 tap leaves an fp32 sum unchanged)
            for (int t = lane; t < 125; t += 32) {
                const int px = x + t / 25 - 2, py = y + (t / 5) % 5 - 2, pz = z + t % 5 - 2;
                float acc = 0.f;
                if (px >= 0 && px < a.h && py >= 0 && py < a.w && pz >= 0 && pz < a.d) {
                    const int mx = px + kH, my = py + kW, mz = pz + kD;
                    const bool min_vol = mx >= 0 && mx < a.h && my >= 0 && my < a.w && mz >= 0 && mz < a.d;
                    const float* fp = a.fix + ((int64_t)px * a.w + py) * a.d + pz;
                    const float* mp = a.mov + (min_vol ? ((int64_t)mx * a.w + my) * a.d + mz : 0);
#pragma unroll 4
                    for (int c = 0; c < a.C; ++c) {
                        const float df = __fsub_rn(__ldg(fp + c * cs), min_vol ? __ldg(mp + c * cs) : 0.f);
                        acc = __fadd_rn(acc, __fmul_rn(df, df));
                    }
                }
                raw[t] = acc;
            }
            __syncwarp();
            if (lane < 27) {
                const int qx = lane / 9 - 1, qy = (lane / 3) % 3 - 1, qz = lane % 3 - 1;
                float b = 0.f;
                if (x + qx >= 0 && x + qx < a.h && y + qy >= 0 && y + qy < a.w && z + qz >= 0 && z + qz < a.d) {
                    float acc = 0.f;
#pragma unroll
                    for (int e = 0; e < 27; ++e) {
                        const int ex = qx + e / 9 - 1, ey = qy + (e / 3) % 3 - 1, ez = qz + e % 3 - 1;
                        acc = __fadd_rn(acc, raw[(ex + 2) * 25 + (ey + 2) * 5 + (ez + 2)]);
                    }
                    b = ca_div27(acc);
                }
                box[lane] = b;
            }
            __syncwarp();
            float acc = 0.f;
#pragma unroll
            for (int e = 0; e < 27; ++e) acc = __fadd_rn(acc, box[e]);
            const float ssd_exact = ca_div27(acc);
            __syncwarp();
            const float d0 = __fsub_rn(a.mesh[lc], s0), d1 = __fsub_rn(a.mesh[L + lc], s1), d2 = __fsub_rn(a.mesh[2 * L + lc], s2);
            const float dist = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
            const unsigned long long k = pack_key(__fadd_rn(ssd_exact, __fmul_rn(a.coeff, dist)), lc);
            best = k < best ? k : best;
        }
        if (lane == 0) s_key[wib] = best;
        __syncthreads();
        if (tid == 0) {
            unsigned long long k = s_key[0];
            for (int i = 1; i < 4; ++i) k = s_key[i] < k ? s_key[i] : k;
            a.keys[v] = k;
        }
        __syncthreads();
    }
}

// certification of the plain argmin (utils.py:87) from the slab minima and runner-ups: the global runner-up is the
// winner slab's own runner-up or another slab's minimum, whichever is smaller
__global__ void __launch_bounds__(256) k_certify_argmin(const float* __restrict__ smin, const float* __restrict__ ssec, int64_t n, int S,
                                                        int* __restrict__ flagged, int* __restrict__ nflagged) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        float best = INFINITY, second = INFINITY;
        int bs = 0;
        for (int sl = 0; sl < S; ++sl) {
            const float m = ca_ld_stream1(smin + (int64_t)sl * n + v);
            second = fminf(second, fmaxf(m, best));
            if (m < best) { best = m; bs = sl; }
        }
        second = fminf(second, ca_ld_stream1(ssec + (int64_t)bs * n + v));
        if (best > 0.f && second <= ca_band_hi(best)) flagged[atomicAdd(nflagged, 1)] = (int)v;
    }
}

// ---- exact pruning of the six passes -----------------------------------------------------------------------------------
// Labels come in blocks of S consecutive ones (same kD and kW, all kH): mesh[1] and mesh[2] are constant inside a block.
// With bmin_b(x) = min of ssd over block b, every label l of the block satisfies, in the reference's own rounded
// arithmetic (every step below is monotone in its arguments),
//     cost_l = RN(ssd_l + RN(coeff * RN(RN(q0 + q1) + q2)))  >=  RN(bmin_b + RN(coeff * RN(q1 + q2)))  =: lb_b
// (q_c = RN(RN(mesh_c - s_c)^2), q0 >= 0).  If lb_b exceeds the cost U of ANY label (here: the previous pass's winner,
// evaluated exactly), no label of block b can be the argmin nor tie with it, so the block is skipped without reading
// it.  The winner's own block, and every block holding an equal cost, has lb <= U and is scanned in label order with
// the same strict `<` as the full pass: the selection is the reference's, bit for bit, ties included.

// one streaming pass over the volume: bmin (S*S, n) block minima, smin (S, n) minima over the S blocks that share kD
// ("slabs" of S*S labels), and the plain argmin (utils.py:87) merged into keys.  blockIdx.y = slab.
// ssec != nullptr (certified mode): also the runner-up cost of every slab, (S, n)
template <int VEC>
__global__ void __launch_bounds__(256) k_blockmin_argmin(const float* __restrict__ ssd, int64_t n, int S, float* __restrict__ bmin,
                                                         float* __restrict__ smin, float* __restrict__ ssec,
                                                         unsigned long long* __restrict__ keys) {
    const int slab = (int)blockIdx.y;
    const int64_t v0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (v0 >= n) return;
    float best[VEC], sm[VEC], sec[VEC];
    int bl[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { best[i] = INFINITY; sm[i] = INFINITY; sec[i] = INFINITY; bl[i] = slab * S * S; }
    const float* sp = ssd + v0;
    for (int b = slab * S; b < (slab + 1) * S; ++b) {
        float bm[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) bm[i] = INFINITY;
#pragma unroll 4
        for (int k = 0; k < S; ++k) {
            const int l = b * S + k;
            float c[VEC];
            if (VEC == 4) {
                const float4 c4 = ca_ld_stream4(reinterpret_cast<const float4*>(sp + (int64_t)l * n));
                c[0] = c4.x; c[1 % VEC] = c4.y; c[2 % VEC] = c4.z; c[3 % VEC] = c4.w;
            } else {
                c[0] = ca_ld_stream1(sp + (int64_t)l * n);
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                bm[i] = fminf(bm[i], c[i]);
                sec[i] = fminf(sec[i], fmaxf(c[i], best[i]));
                if (c[i] < best[i]) { best[i] = c[i]; bl[i] = l; }
            }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) sm[i] = fminf(sm[i], bm[i]);
        if (VEC == 4) *reinterpret_cast<float4*>(bmin + (int64_t)b * n + v0) = make_float4(bm[0], bm[1 % VEC], bm[2 % VEC], bm[3 % VEC]);
        else bmin[(int64_t)b * n + v0] = bm[0];
    }
    if (VEC == 4) *reinterpret_cast<float4*>(smin + (int64_t)slab * n + v0) = make_float4(sm[0], sm[1 % VEC], sm[2 % VEC], sm[3 % VEC]);
    else smin[(int64_t)slab * n + v0] = sm[0];
    if (ssec) {
        if (VEC == 4) *reinterpret_cast<float4*>(ssec + (int64_t)slab * n + v0) = make_float4(sec[0], sec[1 % VEC], sec[2 % VEC], sec[3 % VEC]);
        else ssec[(int64_t)slab * n + v0] = sec[0];
    }
    if (keys) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) atomicMin(keys + v0 + i, pack_key(best[i], bl[i]));
    }
}

// componentwise distance of s to the box [lo, hi], squared and summed in the reference's order: a lower bound of
// RN(RN(q0 + q1) + q2) for every displacement vector inside the box (|RN(m - s)| >= RN(distance), all steps monotone)
static __device__ __forceinline__ float box_dist(const float4& lo, const float4& hi, float s0, float s1, float s2) {
    const float e0 = fmaxf(fmaxf(__fsub_rn(lo.x, s0), __fsub_rn(s0, hi.x)), 0.f);
    const float e1 = fmaxf(fmaxf(__fsub_rn(lo.y, s1), __fsub_rn(s1, hi.y)), 0.f);
    const float e2 = fmaxf(fmaxf(__fsub_rn(lo.z, s2), __fsub_rn(s2, hi.z)), 0.f);
    return __fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2));
}

// bounding boxes of the displacement vectors of every block of S labels and every slab of S blocks:
// boxes = [blo (B) | bhi (B) | slo (S) | shi (S) | axis (S) | range lo, hi | flag] float4.  One CTA; the mesh is a per-call constant.
// axis[k] = (mesh[0, kH=k], mesh[1, kW=k], mesh[2, kD=k]); flag.x = 1.0 when the mesh is the separable grid those
// tables describe (mesh[c, l] depends only on the c-th index of l, bit for bit) -- disp_mesh_t of MIND.py:127 is.
__global__ void __launch_bounds__(256) k_mesh_boxes(const float* __restrict__ mesh, int L, int S, float4* __restrict__ boxes) {
    const int B = S * S;
    float4* blo = boxes;
    float4* bhi = blo + B;
    float4* slo = bhi + B;
    float4* shi = slo + S;
    float4* axis = shi + S;
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    int bad = 0;
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const int kH = l % S, kW = (l / S) % S, kD = l / B;
        bad |= __float_as_uint(mesh[l]) != __float_as_uint(mesh[kH]);
        bad |= __float_as_uint(mesh[L + l]) != __float_as_uint(mesh[L + kW * S]);
        bad |= __float_as_uint(mesh[2 * L + l]) != __float_as_uint(mesh[2 * L + kD * B]);
    }
    if (bad) s_bad = 1;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float4 lo = make_float4(mesh[b * S], mesh[L + b * S], mesh[2 * L + b * S], 0.f), hi = lo;
        for (int k = 1; k < S; ++k) {
            const int l = b * S + k;
            lo.x = fminf(lo.x, mesh[l]); lo.y = fminf(lo.y, mesh[L + l]); lo.z = fminf(lo.z, mesh[2 * L + l]);
            hi.x = fmaxf(hi.x, mesh[l]); hi.y = fmaxf(hi.y, mesh[L + l]); hi.z = fmaxf(hi.z, mesh[2 * L + l]);
        }
        blo[b] = lo; bhi[b] = hi;
    }
    __syncthreads();
    for (int sl = threadIdx.x; sl < S; sl += blockDim.x) {
        float4 lo = blo[sl * S], hi = bhi[sl * S];
        for (int k = 1; k < S; ++k) {
            const float4 a = blo[sl * S + k], c = bhi[sl * S + k];
            lo.x = fminf(lo.x, a.x); lo.y = fminf(lo.y, a.y); lo.z = fminf(lo.z, a.z);
            hi.x = fmaxf(hi.x, c.x); hi.y = fmaxf(hi.y, c.y); hi.z = fmaxf(hi.z, c.z);
        }
        slo[sl] = lo; shi[sl] = hi;
    }
    for (int k = threadIdx.x; k < S; k += blockDim.x) axis[k] = make_float4(mesh[k], mesh[L + k * S], mesh[2 * L + k * B], 0.f);
    __syncthreads();
    if (threadIdx.x == 0) {
        float4 lo = axis[0], hi = axis[0];
        for (int k = 1; k < S; ++k) {
            lo.x = fminf(lo.x, axis[k].x); lo.y = fminf(lo.y, axis[k].y); lo.z = fminf(lo.z, axis[k].z);
            hi.x = fmaxf(hi.x, axis[k].x); hi.y = fmaxf(hi.y, axis[k].y); hi.z = fmaxf(hi.z, axis[k].z);
        }
        axis[S] = lo; axis[S + 1] = hi;
        axis[S + 2] = make_float4(s_bad ? 0.f : 1.f, 0.f, 0.f, 0.f);
    }
}

// one coupled-convex pass that reads the slab minima (S*n floats), the block minima of the slabs that survive, and the
// labels of the blocks that survive.  The label index is uniform across a warp, so mesh[l] is a broadcast load.
// CERT (certified mode, see below): the costs are approximate, so a block is only skipped when its bound exceeds
// U * (1 + 2^-16), the runner-up cost is tracked next to the best one, and a voxel whose runner-up lies within the band of
// its best cost is appended to `flagged` for k_cc_fixup.
template <int VEC, int S, bool CERT>
__global__ void __launch_bounds__(128, 5) k_cost_argmin_pruned(const float* __restrict__ ssd, const float* __restrict__ bmin,
                                                            const float* __restrict__ smin, int64_t n, int L,
                                                            const float* __restrict__ mesh, const float4* __restrict__ boxes,
                                                            const float* __restrict__ soft, float coeff, int row_len,
                                                            const unsigned long long* __restrict__ prev,
                                                            unsigned long long* __restrict__ keys, int* __restrict__ flagged,
                                                            int* __restrict__ nflagged) {
    CA_DYN_SMEM(float4, sbox);
    const int B = S * S;
    if (boxes[2 * B + 3 * S + 2].x != 0.f) return;       // separable mesh: k_cost_argmin_sep does this pass
    for (int i = threadIdx.x; i < 2 * B + 2 * S; i += blockDim.x) sbox[i] = boxes[i];
    __syncthreads();
    const float4* blo = sbox;
    const float4* bhi = blo + B;
    const float4* slo = bhi + B;
    const float4* shi = slo + S;
    // a warp covers 4 rows x 32 voxels (row_len > 0) rather than 128 voxels of one row: neighbours in both directions
    // share their surviving blocks, so the union a warp has to scan is smaller
    int64_t v0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (VEC == 4 && row_len > 0) {
        const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t warp = gt >> 5;
        const int lane = (int)(gt & 31);
        const int tpr = row_len / 32;
        v0 = ((warp / tpr) * 4 + (lane >> 3)) * row_len + (warp % tpr) * 32 + (lane & 7) * 4;
    }
    if (v0 >= n) return;
    float s0[VEC], s1[VEC], s2[VEC], best[VEC], U[VEC], second[VEC];
    int bl[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        best[i] = INFINITY;
        second[i] = INFINITY;
        bl[i] = 0;
        s0[i] = soft[v0 + i]; s1[i] = soft[n + v0 + i]; s2[i] = soft[2 * n + v0 + i];
    }
    // The bound U: the exact cost, under the new soft field and coefficient, of (a) the previous selection and (b) the
    // label next to soft(x).  Any label gives a valid bound.
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        int cand[2];
        cand[0] = (int)(unsigned)(prev[v0 + i] & 0xffffffffull);
        // the label next to soft(x) on the regular displacement grid (mesh[c, l] ~ k_c - hw): one more upper bound
        {
            const float hwf = 0.5f * (float)(S - 1);
            const int kH = ca_clamp((int)rintf(s0[i] + hwf), 0, S - 1), kW = ca_clamp((int)rintf(s1[i] + hwf), 0, S - 1),
                      kD = ca_clamp((int)rintf(s2[i] + hwf), 0, S - 1);
            cand[1] = (kD * S + kW) * S + kH;
        }
        U[i] = INFINITY;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int r = cand[t];
            const float d0 = __fsub_rn(__ldg(mesh + r), s0[i]), d1 = __fsub_rn(__ldg(mesh + L + r), s1[i]), d2 = __fsub_rn(__ldg(mesh + 2 * L + r), s2[i]);
            const float dist = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
            U[i] = fminf(U[i], __fadd_rn(__ldg(ssd + (int64_t)r * n + v0 + i), __fmul_rn(coeff, dist)));
        }
        if (CERT) U[i] = ca_band_hi(U[i]);
    }
    // level 1: which slabs can hold the argmin (S independent loads, no branches)
    unsigned slab_mask = 0;
#pragma unroll
    for (int sl = 0; sl < S; ++sl) {
        float sm[VEC];
        if (VEC == 4) {
            const float4 b4 = ca_ld_stream4(reinterpret_cast<const float4*>(smin + (int64_t)sl * n + v0));
            sm[0] = b4.x; sm[1 % VEC] = b4.y; sm[2 % VEC] = b4.z; sm[3 % VEC] = b4.w;
        } else {
            sm[0] = ca_ld_stream1(smin + (int64_t)sl * n + v0);
        }
        const float4 lo = slo[sl], hi = shi[sl];
        bool need = false;
#pragma unroll
        for (int i = 0; i < VEC; ++i) need = need || !(__fadd_rn(sm[i], __fmul_rn(coeff, box_dist(lo, hi, s0[i], s1[i], s2[i]))) > U[i]);
        slab_mask |= need ? (1u << sl) : 0u;
    }
    const float* sp = ssd + v0;
    for (int sl = 0; sl < S; ++sl) {
        if (!((slab_mask >> sl) & 1u)) continue;
        // level 2: the blocks of this slab
        for (int bb = 0; bb < S; bb += 4) {
            float bmv[4][VEC];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int b = sl * S + ca_min(bb + u, S - 1);
                if (VEC == 4) {
                    const float4 b4 = ca_ld_stream4(reinterpret_cast<const float4*>(bmin + (int64_t)b * n + v0));
                    bmv[u][0] = b4.x; bmv[u][1 % VEC] = b4.y; bmv[u][2 % VEC] = b4.z; bmv[u][3 % VEC] = b4.w;
                } else {
                    bmv[u][0] = ca_ld_stream1(bmin + (int64_t)b * n + v0);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (bb + u >= S) break;
                const int b = sl * S + bb + u;
                const float4 lo = blo[b], hi = bhi[b];
                bool need = false;
#pragma unroll
                for (int i = 0; i < VEC; ++i)
                    need = need || !(__fadd_rn(bmv[u][i], __fmul_rn(coeff, box_dist(lo, hi, s0[i], s1[i], s2[i]))) > U[i]);
                if (!need) continue;
                float cv[S][VEC];                 // the S rows of the block: independent loads, all in flight together
#pragma unroll
                for (int k = 0; k < S; ++k) {
                    const int l = b * S + k;
                    if (VEC == 4) {
                        const float4 c4 = ca_ld_stream4(reinterpret_cast<const float4*>(sp + (int64_t)l * n));
                        cv[k][0] = c4.x; cv[k][1 % VEC] = c4.y; cv[k][2 % VEC] = c4.z; cv[k][3 % VEC] = c4.w;
                    } else {
                        cv[k][0] = ca_ld_stream1(sp + (int64_t)l * n);
                    }
                }
#pragma unroll
                for (int k = 0; k < S; ++k) {
                    const int l = b * S + k;
                    const float* c = cv[k];
                    const float m0 = __ldg(mesh + l), m1 = __ldg(mesh + L + l), m2 = __ldg(mesh + 2 * L + l);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float d0 = __fsub_rn(m0, s0[i]), d1 = __fsub_rn(m1, s1[i]), d2 = __fsub_rn(m2, s2[i]);
                        const float dist = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
                        const float cost = __fadd_rn(c[i], __fmul_rn(coeff, dist));
                        if (CERT) second[i] = fminf(second[i], fmaxf(cost, best[i]));
                        if (cost < best[i]) { best[i] = cost; bl[i] = l; }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        keys[v0 + i] = pack_key(best[i], bl[i]);
        // an exactly zero cost is exact in both arithmetics (every term of the sums is zero), ties included
        if (CERT && best[i] > 0.f && second[i] <= ca_band_hi(best[i])) flagged[atomicAdd(nflagged, 1)] = (int)(v0 + i);
    }
}

// The same pass for the separable displacement grid (the mesh of MIND.py:127): mesh[0,l], mesh[1,l], mesh[2,l] depend
// only on kH, kW, kD, so the squared distances split into per-axis terms that are computed once per slab (kD), once per
// block (kW) and once per label (kH) -- in the reference's rounded operations, so every cost and every bound is the
// value the generic kernel computes, bit for bit, at less than half the instructions.
#ifndef CA_SEP1_MINB
#define CA_SEP1_MINB 8      // resident CTAs the one-voxel form of the selection pass is compiled for (55 registers); 10 / 12 (48 / 40 registers, no spills) measured slower: 2.94 / 2.92 against 2.77 ms per direction
#endif
// VEC consecutive floats as one streaming load
template <int VEC>
static __device__ __forceinline__ void cc_ld_vec(const float* p, float (&o)[VEC]) {
    if (VEC == 4) {
        const float4 t = ca_ld_stream4(reinterpret_cast<const float4*>(p));
        o[0] = t.x; o[1 % VEC] = t.y; o[2 % VEC] = t.z; o[3 % VEC] = t.w;
    } else if (VEC == 2) {
        const float2 t = ca_ld_stream2(reinterpret_cast<const float2*>(p));
        o[0] = t.x; o[1 % VEC] = t.y;
    } else {
        o[0] = ca_ld_stream1(p);
    }
}

template <int VEC, int S, bool CERT, int BG = 4>
__global__ void __launch_bounds__(128, VEC == 4 ? 5 : (VEC == 1 ? CA_SEP1_MINB : 8)) k_cost_argmin_sep(const float* __restrict__ ssd, const float* __restrict__ bmin,
                                                            const float* __restrict__ smin, int64_t n, const float4* __restrict__ boxes,
                                                            const float* __restrict__ soft, float coeff, int row_len, int rpw,
                                                            const unsigned long long* __restrict__ prev,
                                                            unsigned long long* __restrict__ keys, int* __restrict__ flagged,
                                                            int* __restrict__ nflagged) {
    __shared__ float4 sax[S + 2];
    const int B = S * S;
    const float4* tab = boxes + 2 * B + 2 * S;
    if (tab[S + 2].x == 0.f) return;                       // not separable: k_cost_argmin_pruned does this pass
    for (int i = threadIdx.x; i < S + 2; i += blockDim.x) sax[i] = tab[i];
    __syncthreads();
    int64_t v0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (row_len > 0) {      // a warp = rpw rows x (32 / rpw) lanes of VEC voxels: a compact tile prunes alike
        const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const int64_t warp = gt >> 5;
        const int lane = (int)(gt & 31);
        const int lpr = 32 / rpw, tpr = row_len / (lpr * VEC);
        v0 = ((warp / tpr) * rpw + (lane / lpr)) * row_len + (warp % tpr) * (lpr * VEC) + (lane % lpr) * VEC;
    }
    if (v0 >= n) return;
    float s0[VEC], s1[VEC], s2[VEC], best[VEC], U[VEC], second[VEC], r01[VEC];
    int bl[VEC];
    const float4 rlo = sax[S], rhi = sax[S + 1];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        best[i] = INFINITY;
        second[i] = INFINITY;
        bl[i] = 0;
        s0[i] = soft[v0 + i]; s1[i] = soft[n + v0 + i]; s2[i] = soft[2 * n + v0 + i];
        // squared distance to the H and W ranges of the grid: the first two terms of every slab's bound
        const float e0 = fmaxf(fmaxf(__fsub_rn(rlo.x, s0[i]), __fsub_rn(s0[i], rhi.x)), 0.f);
        const float e1 = fmaxf(fmaxf(__fsub_rn(rlo.y, s1[i]), __fsub_rn(s1[i], rhi.y)), 0.f);
        r01[i] = __fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1));
    }
    // The bound U: the exact cost, under the new soft field and coefficient, of (a) the previous selection and (b) the
    // label next to soft(x).  Any label gives a valid bound.
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        int cand[2];
        cand[0] = (int)(unsigned)(prev[v0 + i] & 0xffffffffull);
        {
            const float hwf = 0.5f * (float)(S - 1);
            const int kH = ca_clamp((int)rintf(s0[i] + hwf), 0, S - 1), kW = ca_clamp((int)rintf(s1[i] + hwf), 0, S - 1),
                      kD = ca_clamp((int)rintf(s2[i] + hwf), 0, S - 1);
            cand[1] = (kD * S + kW) * S + kH;
        }
        U[i] = INFINITY;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int r = cand[t];
            const float d0 = __fsub_rn(sax[r % S].x, s0[i]), d1 = __fsub_rn(sax[(r / S) % S].y, s1[i]), d2 = __fsub_rn(sax[r / B].z, s2[i]);
            const float dist = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
            U[i] = fminf(U[i], __fadd_rn(__ldg(ssd + (int64_t)r * n + v0 + i), __fmul_rn(coeff, dist)));
        }
        if (CERT) U[i] = ca_band_hi(U[i]);
    }
    // level 1: which slabs can hold the argmin (S independent loads, no branches)
    unsigned slab_mask = 0;
#pragma unroll
    for (int sl = 0; sl < S; ++sl) {
        float sm[VEC];
        cc_ld_vec<VEC>(smin + (int64_t)sl * n + v0, sm);
        const float a2 = sax[sl].z;
        bool need = false;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float d2 = __fsub_rn(a2, s2[i]);
            need = need || !(__fadd_rn(sm[i], __fmul_rn(coeff, __fadd_rn(r01[i], __fmul_rn(d2, d2)))) > U[i]);
        }
        slab_mask |= need ? (1u << sl) : 0u;
    }
    const float* sp = ssd + v0;
    for (int sl = 0; sl < S; ++sl) {
        if (!((slab_mask >> sl) & 1u)) continue;
        float q2[VEC], e0q[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float d2 = __fsub_rn(sax[sl].z, s2[i]);
            q2[i] = __fmul_rn(d2, d2);
            const float e0 = fmaxf(fmaxf(__fsub_rn(rlo.x, s0[i]), __fsub_rn(s0[i], rhi.x)), 0.f);
            e0q[i] = __fmul_rn(e0, e0);
        }
        // level 2: the blocks of this slab
        for (int bb = 0; bb < S; bb += BG) {          // BG block minima in flight together
            float bmv[BG][VEC];
#pragma unroll
            for (int u = 0; u < BG; ++u) {
                const int b = sl * S + ca_min(bb + u, S - 1);
                cc_ld_vec<VEC>(bmin + (int64_t)b * n + v0, bmv[u]);
            }
#pragma unroll
            for (int u = 0; u < BG; ++u) {
                if (bb + u >= S) break;
                const int b = sl * S + bb + u;
                const float a1 = sax[bb + u].y;
                float q1[VEC];
                bool need = false;
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float d1 = __fsub_rn(a1, s1[i]);
                    q1[i] = __fmul_rn(d1, d1);
                    need = need || !(__fadd_rn(bmv[u][i], __fmul_rn(coeff, __fadd_rn(__fadd_rn(e0q[i], q1[i]), q2[i]))) > U[i]);
                }
                if (!need) continue;
                float cv[S][VEC];                 // the S rows of the block: independent loads, all in flight together
#pragma unroll
                for (int k = 0; k < S; ++k) {
                    const int l = b * S + k;
                    cc_ld_vec<VEC>(sp + (int64_t)l * n, cv[k]);
                }
#pragma unroll
                for (int k = 0; k < S; ++k) {
                    const int l = b * S + k;
                    const float a0 = sax[k].x;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float d0 = __fsub_rn(a0, s0[i]);
                        const float dist = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), q1[i]), q2[i]);
                        const float cost = __fadd_rn(cv[k][i], __fmul_rn(coeff, dist));
                        if (CERT) second[i] = fminf(second[i], fmaxf(cost, best[i]));
                        if (cost < best[i]) { best[i] = cost; bl[i] = l; }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        keys[v0 + i] = pack_key(best[i], bl[i]);
        if (CERT && best[i] > 0.f && second[i] <= ca_band_hi(best[i])) flagged[atomicAdd(nflagged, 1)] = (int)(v0 + i);
    }
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

static int g_cc_prune = -1;   // CA_CC_PRUNE=0 disables the exact pruning (A/B timing of the plain streaming passes)
static bool cc_prune_enabled() {
    if (g_cc_prune < 0) { const char* e = getenv("CA_CC_PRUNE"); g_cc_prune = (e && atoi(e) == 0) ? 0 : 1; }
    return g_cc_prune == 1;
}

// bmin (S*S,n) followed by smin (S,n) in one buffer, and, when keys != nullptr, the plain argmin as keys
int ca_launch_blockmin_argmin(const float* ssd, int64_t n, int S, float* bmin, unsigned long long* keys, cudaStream_t st,
                              float* ssec = nullptr) {
    float* smin = bmin + (int64_t)S * S * n;
    const int vec = (n % 4 == 0 && ((((uintptr_t)ssd) | ((uintptr_t)bmin) | ((uintptr_t)ssec)) & 15) == 0) ? 4 : 1;
    const unsigned gx = (unsigned)ca_divup(ca_divup(n, vec), 256);
    if (keys && cudaMemsetAsync(keys, 0xff, (size_t)n * 8, st) != cudaSuccess) {
        ca_set_error("blockmin_argmin: memset failed");
        return CA_ERR_LAUNCH;
    }
    const dim3 grid(gx, (unsigned)S), block(256);
    void* tm = ca_timing_begin(CA_T_BLOCKMIN, st);
    if (vec == 4) { auto k = k_blockmin_argmin<4>; CA_LAUNCH(k, grid, block, 0, st, ssd, n, S, bmin, smin, ssec, keys); }
    else { auto k = k_blockmin_argmin<1>; CA_LAUNCH(k, grid, block, 0, st, ssd, n, S, bmin, smin, ssec, keys); }
    ca_timing_end(tm, st);
    CA_CHECK_LAUNCH("blockmin_argmin");
    return CA_OK;
}

// boxes of the mesh for the pruned passes: 2*S*S + 2*S float4 placed after bmin/smin in the same buffer
static float4* cc_boxes_of(float* bmin, int64_t n, int S) {
    return (float4*)(((uintptr_t)(bmin + (int64_t)(S * S + S) * n) + 255) & ~(uintptr_t)255);
}
static size_t cc_bmin_bytes(int64_t n, int S) { return ((size_t)S * S + S) * n * 4 + 256 + (size_t)(2 * S * S + 3 * S + 3) * 16; }

int ca_launch_mesh_boxes(const float* mesh, int S, float4* boxes, cudaStream_t st) {
    CA_LAUNCH(k_mesh_boxes, dim3(1), dim3(256), 0, st, mesh, S * S * S, S, boxes);
    CA_CHECK_LAUNCH("mesh_boxes");
    return CA_OK;
}

// one pruned pass; keys holds the previous selection on entry and the new one on exit (each thread reads its own
// entries before it writes them).  flagged != nullptr: certified mode (approximate costs, see ca_band_hi).
int ca_launch_cost_argmin_pruned(const float* ssd, const float* bmin, int64_t n, int S, const float* mesh, const float4* boxes,
                                 const float* soft, float coeff, int row_len, unsigned long long* keys, cudaStream_t st,
                                 int* flagged = nullptr, int* nflagged = nullptr) {
    if (row_len <= 0 || row_len % 32 != 0 || (n / row_len) % 4 != 0 || n % row_len != 0) row_len = 0;   // 4x32 warp tiles need whole tiles
    const int L = S * S * S, B = S * S;
    const float* smin = bmin + (int64_t)B * n;
    const int vec = (n % 4 == 0 && ((((uintptr_t)ssd) | ((uintptr_t)bmin)) & 15) == 0) ? 4 : 1;
    const unsigned gx = (unsigned)ca_divup(ca_divup(n, vec), 128);
    const size_t smem = ((size_t)2 * B + 2 * S) * sizeof(float4);
    const bool cert = flagged != nullptr;
    void* tm = ca_timing_begin(CA_T_SELECT, st);
#define CA_PRUNED_K(SS, VV, CC) { auto k = k_cost_argmin_pruned<VV, SS, CC>; CA_LAUNCH(k, dim3(gx), dim3(128), smem, st, ssd, bmin, smin, n, L, mesh, boxes, soft, coeff, row_len, keys, keys, flagged, nflagged); }
#define CA_PRUNED(SS)                                           \
    case SS:                                                    \
        if (vec == 4 && cert) CA_PRUNED_K(SS, 4, true)          \
        else if (vec == 4) CA_PRUNED_K(SS, 4, false)            \
        else if (cert) CA_PRUNED_K(SS, 1, true)                 \
        else CA_PRUNED_K(SS, 1, false)                          \
        break;
    switch (S) {
        CA_PRUNED(3) CA_PRUNED(5) CA_PRUNED(7) CA_PRUNED(9) CA_PRUNED(11) CA_PRUNED(13) CA_PRUNED(15) CA_PRUNED(17)
        default: ca_set_error("cost_argmin_pruned: disp_hw out of range"); return CA_ERR_UNSUPPORTED;
    }
#undef CA_PRUNED
#undef CA_PRUNED_K
    // Voxels per thread of the separable kernel.  The pass is latency-bound (dependent loads slab -> block -> labels) and a warp
    // evaluates every block ANY of its voxels needs: fewer voxels per thread = more warps in flight and fewer voxels sharing a
    // pruning decision.  Measured per direction at config 2 (coupled_certified, six passes + block minima): 4 voxels 3.43 ms,
    // 2 voxels 3.11 ms, 1 voxel 2.87 ms (warp = 32 voxels of a row) / 2.77 ms (warp = 2 rows x 16 voxels, the default);
    // CA_CC_VEC = 1 | 2 | 4, CA_CC_RPW = rows of the 1-voxel form's warp tile.
    static int sep_vec_env = -1;
    if (sep_vec_env < 0) { const char* e = getenv("CA_CC_VEC"); sep_vec_env = e && *e ? atoi(e) : 1; }
    const int svec = (vec == 4 && (sep_vec_env == 2 || sep_vec_env == 1)) ? sep_vec_env : vec;
    const unsigned gxs = (unsigned)ca_divup(ca_divup(n, svec), 128);
    // rows of a warp's tile: 4 for the 4- and 2-voxel forms (as before), CA_CC_RPW for the 1-voxel form (default 2 rows x 16 voxels)
    static int sep_rpw_env = -1;
    if (sep_rpw_env < 0) { const char* e = getenv("CA_CC_RPW"); sep_rpw_env = e && *e ? atoi(e) : 2; }
    int rpw = svec == 1 ? sep_rpw_env : 4;
    if (rpw != 1 && rpw != 2 && rpw != 4 && rpw != 8) rpw = 1;
    // whole tiles only: rows of a multiple of (32 / rpw) * svec voxels, a multiple of rpw rows
    const int srow = (row_len > 0 && row_len % ((32 / rpw) * svec) == 0 && (n / row_len) % rpw == 0) ? row_len : 0;
    static int sep_bg_env = -1;
    if (sep_bg_env < 0) { const char* e = getenv("CA_CC_BG"); sep_bg_env = e && *e ? atoi(e) : 0; }
#define CA_SEP_KB(SS, VV, CC) { auto k = k_cost_argmin_sep<VV, SS, CC, SS>; CA_LAUNCH(k, dim3(gxs), dim3(128), 0, st, ssd, bmin, smin, n, boxes, soft, coeff, srow, rpw, keys, keys, flagged, nflagged); }
#define CA_SEP_K(SS, VV, CC) { auto k = k_cost_argmin_sep<VV, SS, CC>; CA_LAUNCH(k, dim3(gxs), dim3(128), 0, st, ssd, bmin, smin, n, boxes, soft, coeff, srow, rpw, keys, keys, flagged, nflagged); }
#define CA_SEP(SS)                                              \
    case SS:                                                    \
        if (svec == 2 && cert && sep_bg_env == 1) CA_SEP_KB(SS, 2, true)  \
        else if (svec == 2 && cert) CA_SEP_K(SS, 2, true)       \
        else if (svec == 2) CA_SEP_K(SS, 2, false)              \
        else if (svec == 4 && cert) CA_SEP_K(SS, 4, true)       \
        else if (svec == 4) CA_SEP_K(SS, 4, false)              \
        else if (cert) CA_SEP_K(SS, 1, true)                    \
        else CA_SEP_K(SS, 1, false)                             \
        break;
    switch (S) {      // the kernel for a separable mesh; exactly one of the two does the work (flag written by k_mesh_boxes)
        CA_SEP(3) CA_SEP(5) CA_SEP(7) CA_SEP(9) CA_SEP(11) CA_SEP(13) CA_SEP(15) CA_SEP(17)
        default: break;
    }
#undef CA_SEP
#undef CA_SEP_K
#undef CA_SEP_KB
    ca_timing_end(tm, st);
    CA_CHECK_LAUNCH("cost_argmin_pruned");
    return CA_OK;
}

extern "C" size_t ca_coupled_convex_workspace_bytes(int h, int w, int d, int disp_hw) {
    const size_t n = (size_t)h * w * d, S = 2 * (size_t)disp_hw + 1;
    return n * 8 + 256 + cc_bmin_bytes((int64_t)n, (int)S) + 256;
}

extern "C" int ca_coupled_convex(const float* ssd, const int64_t* argmin, const float* mesh, int h, int w, int d,
                                 int disp_hw, float* disp_soft, int64_t* sel_out, void* workspace, size_t workspace_bytes,
                                 ca_stream_t stream) {
    CA_CHECK_ARG(ssd && mesh && disp_soft && workspace, "coupled_convex: null pointer");
    CA_CHECK_ARG(h > 0 && w > 0 && d > 0 && disp_hw >= 1 && disp_hw <= 8 && h < 65536 && w < 65536, "coupled_convex: bad dims");
    if (workspace_bytes < ca_coupled_convex_workspace_bytes(h, w, d, disp_hw)) {
        ca_set_error("coupled_convex: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    static const float coeffs[6] = {0.003f, 0.01f, 0.03f, 0.1f, 0.3f, 1.f};   // utils.py:98
    cudaStream_t st = (cudaStream_t)stream;
    const int S = 2 * disp_hw + 1, L = S * S * S;
    const int64_t n = (int64_t)h * w * d;
    unsigned long long* keys = (unsigned long long*)workspace;
    float* bmin = (float*)(((uintptr_t)(keys + n) + 255) & ~(uintptr_t)255);
    const bool prune = cc_prune_enabled();
    int rc;
    // argmin == NULL: the plain argmin of utils.py:87 is taken here, in the same pass that builds the block minima
    if (argmin) CA_LAUNCH(k_i64_to_keys, dim3(small_grid(n)), dim3(256), 0, st, argmin, n, L, keys);
    float4* boxes = cc_boxes_of(bmin, n, S);
    if (prune) {
        if ((rc = ca_launch_blockmin_argmin(ssd, n, S, bmin, argmin ? nullptr : keys, st)) != CA_OK) return rc;
        if ((rc = ca_launch_mesh_boxes(mesh, S, boxes, st)) != CA_OK) return rc;
    } else if (!argmin) {
        if ((rc = ca_launch_cost_argmin(ssd, n, L, nullptr, nullptr, 0.f, keys, st)) != CA_OK) return rc;
    }
    CA_LAUNCH(k_soft_from_keys, dim3((unsigned)ca_divup(d, 128), (unsigned)w, (unsigned)h), dim3(128), 0, st, keys, mesh, h, w, d, L, disp_soft);
    CA_CHECK_LAUNCH("coupled_convex/init");
    for (int j = 0; j < 6; ++j) {
        if (prune) rc = ca_launch_cost_argmin_pruned(ssd, bmin, n, S, mesh, boxes, disp_soft, coeffs[j], d, keys, st);
        else rc = ca_launch_cost_argmin(ssd, n, L, mesh, disp_soft, coeffs[j], keys, st);
        if (rc != CA_OK) return rc;
        if (sel_out) {
            rc = ca_launch_keys_to_i64(keys, n, sel_out + (int64_t)j * n, st);
            if (rc != CA_OK) return rc;
        }
        CA_LAUNCH(k_soft_from_keys, dim3((unsigned)ca_divup(d, 128), (unsigned)w, (unsigned)h), dim3(128), 0, st, keys, mesh, h, w, d, L, disp_soft);
        CA_CHECK_LAUNCH("coupled_convex/soft");
    }
    return CA_OK;
}

// ---- coupled_convex on the approximate volume of ca_correlate_fast, selections certified exact -------------------------
// workspace: [keys n u64 | bmin,smin,boxes | flagged n i32 | counters 8 i32]
extern "C" size_t ca_coupled_convex_certified_workspace_bytes(int h, int w, int d, int disp_hw) {
    const size_t n = (size_t)h * w * d;
    const size_t S = 2 * (size_t)disp_hw + 1;
    return ca_coupled_convex_workspace_bytes(h, w, d, disp_hw) + n * 4 + 256 + 64 + 256 + S * n * 4 + 256;
}

extern "C" int ca_coupled_convex_certified(const float* ssd, const float* fix, const float* mov, int C, const float* mesh, int h,
                                           int w, int d, int disp_hw, float* disp_soft, int64_t* sel_out, int32_t* flagged_out,
                                           void* workspace, size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(ssd && fix && mov && mesh && disp_soft && workspace, "coupled_convex_certified: null pointer");
    CA_CHECK_ARG(C > 0 && h > 0 && w > 0 && d > 0 && disp_hw >= 1 && disp_hw <= 8 && h < 65536 && w < 65536, "coupled_convex_certified: bad dims");
    CA_CHECK_ARG((int64_t)h * w * d < (1ll << 31), "coupled_convex_certified: coarse grid too large");
    if (workspace_bytes < ca_coupled_convex_certified_workspace_bytes(h, w, d, disp_hw)) {
        ca_set_error("coupled_convex_certified: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    static const float coeffs[6] = {0.003f, 0.01f, 0.03f, 0.1f, 0.3f, 1.f};   // utils.py:98
    cudaStream_t st = (cudaStream_t)stream;
    const int S = 2 * disp_hw + 1, L = S * S * S;
    const int64_t n = (int64_t)h * w * d;
    unsigned long long* keys = (unsigned long long*)workspace;
    float* bmin = (float*)(((uintptr_t)(keys + n) + 255) & ~(uintptr_t)255);
    float4* boxes = cc_boxes_of(bmin, n, S);
    int* flagged = (int*)(((uintptr_t)((char*)workspace + ca_coupled_convex_workspace_bytes(h, w, d, disp_hw)) + 255) & ~(uintptr_t)255);
    int* counters = (int*)(((uintptr_t)(flagged + n) + 63) & ~(uintptr_t)63);
    float* ssec = (float*)(((uintptr_t)(counters + 16) + 255) & ~(uintptr_t)255);
    int rc;
    if (cudaMemsetAsync(counters, 0, 8 * sizeof(int), st) != cudaSuccess || cudaMemsetAsync(disp_soft, 0, (size_t)3 * n * 4, st) != cudaSuccess) {
        ca_set_error("coupled_convex_certified: memset failed");
        return CA_ERR_LAUNCH;
    }
    // approximate argmin, block/slab minima and slab runner-ups in one streaming pass; k_certify_argmin then flags the
    // voxels whose runner-up lies within the band (the fix-up runs with coeff = 0, soft = 0: regulariser exactly +0)
    if ((rc = ca_launch_blockmin_argmin(ssd, n, S, bmin, keys, st, ssec)) != CA_OK) return rc;
    if ((rc = ca_launch_mesh_boxes(mesh, S, boxes, st)) != CA_OK) return rc;
    FixupArgs fa;
    fa.ssd = ssd; fa.mesh = mesh; fa.soft = disp_soft; fa.fix = fix; fa.mov = mov; fa.flagged = flagged; fa.keys = keys;
    fa.n = n; fa.C = C; fa.h = h; fa.w = w; fa.d = d; fa.hw = disp_hw; fa.x0 = 0;
    for (int j = -1; j < 6; ++j) {
        const float coeff = j < 0 ? 0.f : coeffs[j];
        int* cnt = counters + (j + 1);
        if (j < 0) {
            CA_LAUNCH(k_certify_argmin, dim3(small_grid(n)), dim3(256), 0, st, bmin + (int64_t)S * S * n, ssec, n, S, flagged, cnt);
            CA_CHECK_LAUNCH("coupled_convex_certified/certify_argmin");
        } else if ((rc = ca_launch_cost_argmin_pruned(ssd, bmin, n, S, mesh, boxes, disp_soft, coeff, d, keys, st, flagged, cnt)) != CA_OK) {
            return rc;
        }
        fa.coeff = coeff; fa.nflagged = cnt;
        void* tmf = ca_timing_begin(CA_T_FIXUP, st);
        CA_LAUNCH(k_cc_fixup, dim3(148 * 8), dim3(128), 0, st, fa);
        ca_timing_end(tmf, st);
        CA_CHECK_LAUNCH("coupled_convex_certified/fixup");
        if (sel_out) {
            rc = ca_launch_keys_to_i64(keys, n, sel_out + (int64_t)(j + 1) * n, st);
            if (rc != CA_OK) return rc;
        }
        CA_LAUNCH(k_soft_from_keys, dim3((unsigned)ca_divup(d, 128), (unsigned)w, (unsigned)h), dim3(128), 0, st, keys, mesh, h, w, d, L, disp_soft);
        CA_CHECK_LAUNCH("coupled_convex_certified/soft");
    }
    if (flagged_out && cudaMemcpyAsync(flagged_out, counters, 7 * sizeof(int), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        ca_set_error("coupled_convex_certified: copy failed");
        return CA_ERR_LAUNCH;
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

extern "C" size_t ca_cc_select_workspace_bytes(int64_t n) { return (size_t)n * 8 + 256 + (2 * 17 * 17 + 3 * 17 + 3) * 16 + 256; }

// block minima (S*S, n) of a slab's cost volume and, when argmin != NULL, its plain argmin (utils.py:87) in the same pass
extern "C" int ca_cc_blockmin(const float* ssd, int64_t n, int disp_hw, float* bmin, int64_t* argmin, void* workspace,
                              size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(ssd && bmin && workspace && n > 0 && disp_hw >= 1 && disp_hw <= 8, "cc_blockmin: bad argument");
    if (workspace_bytes < ca_cc_select_workspace_bytes(n)) {
        ca_set_error("cc_blockmin: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    unsigned long long* keys = (unsigned long long*)workspace;
    int rc = ca_launch_blockmin_argmin(ssd, n, 2 * disp_hw + 1, bmin, argmin ? keys : nullptr, (cudaStream_t)stream);
    if (rc != CA_OK || !argmin) return rc;
    return ca_launch_keys_to_i64(keys, n, argmin, (cudaStream_t)stream);
}

extern "C" int ca_cc_select(const float* ssd, int64_t n, int disp_hw, const float* mesh, const float* soft, float coeff,
                            const float* bmin, const int64_t* prev, int64_t* sel, void* workspace, size_t workspace_bytes,
                            ca_stream_t stream) {
    CA_CHECK_ARG(ssd && sel && workspace && n > 0 && disp_hw >= 1 && disp_hw <= 8, "cc_select: bad argument");
    CA_CHECK_ARG((mesh == nullptr) == (soft == nullptr), "cc_select: mesh and soft go together");
    CA_CHECK_ARG((bmin == nullptr) == (prev == nullptr) && (!bmin || mesh), "cc_select: bmin, prev (and mesh, soft) go together");
    if (workspace_bytes < ca_cc_select_workspace_bytes(n)) {
        ca_set_error("cc_select: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int S = 2 * disp_hw + 1;
    unsigned long long* keys = (unsigned long long*)workspace;
    int rc;
    if (bmin && cc_prune_enabled()) {
        CA_LAUNCH(k_i64_to_keys, dim3(small_grid(n)), dim3(256), 0, st, prev, n, S * S * S, keys);
        CA_CHECK_LAUNCH("cc_select/keys");
        float4* boxes = (float4*)(((uintptr_t)(keys + n) + 255) & ~(uintptr_t)255);
        if ((rc = ca_launch_mesh_boxes(mesh, S, boxes, st)) != CA_OK) return rc;
        rc = ca_launch_cost_argmin_pruned(ssd, bmin, n, S, mesh, boxes, soft, coeff, 0, keys, st);
    } else {
        rc = ca_launch_cost_argmin(ssd, n, S * S * S, mesh, soft, coeff, keys, st);
    }
    if (rc != CA_OK) return rc;
    return ca_launch_keys_to_i64(keys, n, sel, st);
}

// ---- the certified stage in pieces (slab-parallel path): the same kernels restricted to the H planes [x_begin, x_end) ----
// workspace: [keys n u64 | boxes | flagged n i32 | counter | zero soft 3n f32]
extern "C" size_t ca_cc_certified_workspace_bytes(int64_t n) {
    return ca_cc_select_workspace_bytes(n) + (size_t)n * 4 + 256 + 256 + (size_t)3 * n * 4 + 256;
}

namespace {
struct CertWs { unsigned long long* keys; float4* boxes; int* flagged; int* counter; float* zsoft; };
CertWs cert_ws(void* workspace, int64_t n) {
    CertWs c;
    c.keys = (unsigned long long*)workspace;
    c.boxes = (float4*)(((uintptr_t)(c.keys + n) + 255) & ~(uintptr_t)255);
    c.flagged = (int*)(((uintptr_t)((char*)workspace + ca_cc_select_workspace_bytes(n)) + 255) & ~(uintptr_t)255);
    c.counter = (int*)(((uintptr_t)(c.flagged + n) + 255) & ~(uintptr_t)255);
    c.zsoft = (float*)(((uintptr_t)(c.counter + 16) + 255) & ~(uintptr_t)255);
    return c;
}
}  // namespace

// bmin: (S*S + 2*S, n) -- block minima, slab minima, slab runner-ups; argmin (n) int64: the plain argmin of utils.py:87,
// certified exact (fix, mov: the whole (C,h,w,d) feature volumes the slab's approximate cost volume was built from)
extern "C" int ca_cc_blockmin_certified(const float* ssd, int64_t n, int disp_hw, const float* mesh, float* bmin, int64_t* argmin,
                                        const float* fix, const float* mov, int C, int h, int w, int d, int x_begin, void* workspace,
                                        size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(ssd && mesh && bmin && argmin && fix && mov && workspace && n > 0 && disp_hw >= 1 && disp_hw <= 8, "cc_blockmin_certified: bad argument");
    CA_CHECK_ARG(C > 0 && h > 0 && w > 0 && d > 0 && x_begin >= 0 && n % ((int64_t)w * d) == 0 && x_begin + n / ((int64_t)w * d) <= h,
                 "cc_blockmin_certified: the slab does not fit the volume");
    if (workspace_bytes < ca_cc_certified_workspace_bytes(n)) {
        ca_set_error("cc_blockmin_certified: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int S = 2 * disp_hw + 1;
    const CertWs c = cert_ws(workspace, n);
    float* smin = bmin + (int64_t)S * S * n;
    float* ssec = smin + (int64_t)S * n;
    if (cudaMemsetAsync(c.counter, 0, 16 * sizeof(int), st) != cudaSuccess || cudaMemsetAsync(c.zsoft, 0, (size_t)3 * n * 4, st) != cudaSuccess) {
        ca_set_error("cc_blockmin_certified: memset failed");
        return CA_ERR_LAUNCH;
    }
    int rc = ca_launch_blockmin_argmin(ssd, n, S, bmin, c.keys, st, ssec);
    if (rc != CA_OK) return rc;
    CA_LAUNCH(k_certify_argmin, dim3(small_grid(n)), dim3(256), 0, st, smin, ssec, n, S, c.flagged, c.counter);
    FixupArgs fa;
    fa.ssd = ssd; fa.mesh = mesh; fa.soft = c.zsoft; fa.fix = fix; fa.mov = mov; fa.flagged = c.flagged; fa.nflagged = c.counter;
    fa.keys = c.keys; fa.n = n; fa.coeff = 0.f; fa.C = C; fa.h = h; fa.w = w; fa.d = d; fa.hw = disp_hw; fa.x0 = x_begin;
    CA_LAUNCH(k_cc_fixup, dim3(148 * 8), dim3(128), 0, st, fa);
    CA_CHECK_LAUNCH("cc_blockmin_certified");
    return ca_launch_keys_to_i64(c.keys, n, argmin, st);
}

// one certified selection pass over a slab (bmin from ca_cc_blockmin_certified, prev = the previous selection)
extern "C" int ca_cc_select_certified(const float* ssd, int64_t n, int disp_hw, const float* mesh, const float* soft, float coeff,
                                      const float* bmin, const int64_t* prev, int64_t* sel, const float* fix, const float* mov, int C,
                                      int h, int w, int d, int x_begin, void* workspace, size_t workspace_bytes, ca_stream_t stream) {
    CA_CHECK_ARG(ssd && mesh && soft && bmin && prev && sel && fix && mov && workspace && n > 0 && disp_hw >= 1 && disp_hw <= 8,
                 "cc_select_certified: bad argument");
    CA_CHECK_ARG(C > 0 && h > 0 && w > 0 && d > 0 && x_begin >= 0 && n % ((int64_t)w * d) == 0 && x_begin + n / ((int64_t)w * d) <= h,
                 "cc_select_certified: the slab does not fit the volume");
    if (workspace_bytes < ca_cc_certified_workspace_bytes(n)) {
        ca_set_error("cc_select_certified: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int S = 2 * disp_hw + 1;
    const CertWs c = cert_ws(workspace, n);
    if (cudaMemsetAsync(c.counter, 0, 16 * sizeof(int), st) != cudaSuccess) {
        ca_set_error("cc_select_certified: memset failed");
        return CA_ERR_LAUNCH;
    }
    CA_LAUNCH(k_i64_to_keys, dim3(small_grid(n)), dim3(256), 0, st, prev, n, S * S * S, c.keys);
    CA_CHECK_LAUNCH("cc_select_certified/keys");
    int rc;
    if ((rc = ca_launch_mesh_boxes(mesh, S, c.boxes, st)) != CA_OK) return rc;
    if ((rc = ca_launch_cost_argmin_pruned(ssd, bmin, n, S, mesh, c.boxes, soft, coeff, 0, c.keys, st, c.flagged, c.counter)) != CA_OK) return rc;
    FixupArgs fa;
    fa.ssd = ssd; fa.mesh = mesh; fa.soft = soft; fa.fix = fix; fa.mov = mov; fa.flagged = c.flagged; fa.nflagged = c.counter;
    fa.keys = c.keys; fa.n = n; fa.coeff = coeff; fa.C = C; fa.h = h; fa.w = w; fa.d = d; fa.hw = disp_hw; fa.x0 = x_begin;
    CA_LAUNCH(k_cc_fixup, dim3(148 * 8), dim3(128), 0, st, fa);
    CA_CHECK_LAUNCH("cc_select_certified/fixup");
    return ca_launch_keys_to_i64(c.keys, n, sel, st);
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
