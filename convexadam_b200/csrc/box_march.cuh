// Shared pieces of the marching 3x3x3 box filters (correlate.cu: both cost-volume filters fused; adam.cu: the B^3
// smoothing of the displacement field).
//
// F.avg_pool3d(x, 3, stride=1, padding=1) sums the in-volume taps in ascending (dH,dW,dD) order into one fp32
// accumulator and divides by 27.  A thread that marches along H with three running sums (the outputs x-1, x, x+1 that
// plane x contributes to) reproduces exactly that order while reading every input value once per plane: taps that
// lie outside the volume are +0 and leave the accumulator unchanged, bit for bit.
#pragma once
#include "ca_platform.h"

// acc <- acc + row taps (dD = -1, 0, +1) for 4 consecutive voxels; FRESH: the accumulator starts here (0 + t == t)
template <bool FRESH>
static __device__ __forceinline__ void box_add_row(float (&a)[4], float hl, const float4& v, float hr) {
    if (FRESH) {
        a[0] = __fadd_rn(__fadd_rn(hl, v.x), v.y);
        a[1] = __fadd_rn(__fadd_rn(v.x, v.y), v.z);
        a[2] = __fadd_rn(__fadd_rn(v.y, v.z), v.w);
        a[3] = __fadd_rn(__fadd_rn(v.z, v.w), hr);
    } else {
        a[0] = __fadd_rn(__fadd_rn(__fadd_rn(a[0], hl), v.x), v.y);
        a[1] = __fadd_rn(__fadd_rn(__fadd_rn(a[1], v.x), v.y), v.z);
        a[2] = __fadd_rn(__fadd_rn(__fadd_rn(a[2], v.y), v.z), v.w);
        a[3] = __fadd_rn(__fadd_rn(__fadd_rn(a[3], v.z), v.w), hr);
    }
}

// Plane X (4 rows of it: y0-1 .. y0+2 for the thread's two output rows y0, y0+1) enters the three running sums:
// set A = R starts output X+1 (its dH=-1 taps), set B continues output X, set C finishes output X-1.
template <int R>
static __device__ __forceinline__ void box_add_plane(float (&s)[3][2][4], const float4 (&v)[4], const float (&hl)[4],
                                                     const float (&hr)[4]) {
    constexpr int A = R, B = (R + 2) % 3, Cs = (R + 1) % 3;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const int dy = j - o - 1;
            if (dy < -1 || dy > 1) continue;
            if (dy == -1) box_add_row<true>(s[A][o], hl[j], v[j], hr[j]);
            else box_add_row<false>(s[A][o], hl[j], v[j], hr[j]);
            box_add_row<false>(s[B][o], hl[j], v[j], hr[j]);
            box_add_row<false>(s[Cs][o], hl[j], v[j], hr[j]);
        }
}

struct BoxLane {
    bool first, last;       // first / last quad of a row: the neighbour is outside the volume (zero)
    bool fb_l, fb_r;        // the neighbour quad lives in another warp: read the value from memory instead
};
// the D-neighbours of a quad: lane-1 holds z0-4..z0-1, lane+1 holds z0+4..z0+7 (threads are z-fastest)
static __device__ __forceinline__ void box_halo(const float4& v, const float* quad, const BoxLane& bl, float& hl, float& hr) {
    hl = __shfl_up_sync(0xffffffffu, v.w, 1);
    hr = __shfl_down_sync(0xffffffffu, v.x, 1);
    if (bl.fb_l) hl = quad[-1];
    if (bl.fb_r) hr = quad[4];
    if (bl.first) hl = 0.f;
    if (bl.last) hr = 0.f;
}

