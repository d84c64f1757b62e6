// The `use_mask` pre-processing of extract_features (reference src/convexAdam/convex_adam_MIND.py:36-54) on the device:
//
//   k_mask_erode   (avg3(mask) > 0.9) with avg3 = ReplicationPad3d(1) + AvgPool3d(3, stride 1)             MIND.py:38-42
//   k_ft_line<AX>  scipy.ndimage.distance_transform_edt(mask_sub == 0, return_indices=True)               MIND.py:43
//                  = for every voxel the index of the NEAREST voxel inside the (eroded, 2x subsampled) mask
//
// The reference runs the distance transform on the host (scipy) in the middle of the GPU path.  scipy's feature transform
// is Maurer's algorithm (ni_morphology.c: _ComputeFT / _VoronoiFT): axis by axis, every line keeps the lower envelope of
// the parabolas of the features found so far.  The kernels below restate it line for line -- same axis order (0, 1, 2),
// same candidate-removal test and same `<=` in the query walk -- because among EQUIDISTANT inside voxels the choice is
// decided by exactly those comparisons, and the chosen voxel's intensity is what fills the outside of the mask.  One
// thread per line; coordinates are small integers, so the double arithmetic of the tests is exact.
// The gather, the 2x trilinear upsampling (ca_trilinear) and the final select stay with the caller (convex_adam_MIND.py).
#include "ca_platform.h"

__global__ void __launch_bounds__(256) k_mask_erode(const float* __restrict__ mask, int H, int W, int D, float* __restrict__ out) {
    const int64_t n = (int64_t)H * W * D;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(o % D), y = (int)((o / D) % W), x = (int)(o / ((int64_t)D * W));
        float s = 0.f;                                   // torch avg_pool3d: ascending (dH, dW, dD) sum, then a true division
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz) {
                    const int xx = ca_clamp(x + dx, 0, H - 1), yy = ca_clamp(y + dy, 0, W - 1), zz = ca_clamp(z + dz, 0, D - 1);   // replication pad
                    s = __fadd_rn(s, __ldg(mask + ((int64_t)xx * W + yy) * D + zz));
                }
        out[o] = ca_div(s, 27.f) > 0.9f ? 1.f : 0.f;
    }
}

// F: (3, h, w, d) int32 feature coordinates (F[0] < 0: no feature known yet).  AX = the axis the lines run along.
// scratch: 4 ints per voxel (3 feature coordinates + the envelope stack), a private segment per line.
template <int AX>
__global__ void __launch_bounds__(128) k_ft_line(const float* __restrict__ inside, int* __restrict__ F, int h, int w, int d,
                                                 int* __restrict__ scratch) {
    const int dims[3] = {h, w, d};
    const int len = dims[AX];
    const int o1 = AX == 0 ? 1 : 0, o2 = AX == 2 ? 1 : 2;            // the two other axes
    const int64_t nlines = (int64_t)dims[o1] * dims[o2];
    const int64_t n = (int64_t)h * w * d;
    const int64_t strides[3] = {(int64_t)w * d, d, 1};
    for (int64_t line = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; line < nlines; line += (int64_t)gridDim.x * blockDim.x) {
        int coor[3];
        coor[o1] = (int)(line / dims[o2]);
        coor[o2] = (int)(line % dims[o2]);
        coor[AX] = 0;
        const int64_t base = coor[o1] * strides[o1] + coor[o2] * strides[o2];
        int* f = scratch + line * len * 4;              // f[ii*3 + jj]
        int* g = f + (int64_t)len * 3;
        for (int ii = 0; ii < len; ++ii) {
            const int64_t v = base + ii * strides[AX];
            if (AX == 0) {
                // _ComputeFT, d == 0: a voxel of the set is its own feature, the others have none yet
                const bool in = inside[v] != 0.f;
                f[ii * 3 + 0] = in ? ii : -1;
                f[ii * 3 + 1] = coor[1];
                f[ii * 3 + 2] = coor[2];
            } else {
                f[ii * 3 + 0] = F[v]; f[ii * 3 + 1] = F[n + v]; f[ii * 3 + 2] = F[2 * n + v];
            }
        }
        // _VoronoiFT: build the lower envelope
        int l = -1;
        for (int ii = 0; ii < len; ++ii) {
            if (f[ii * 3] < 0) continue;
            const double fd = f[ii * 3 + AX];
            double wR = 0.0;
            for (int jj = 0; jj < 3; ++jj)
                if (jj != AX) { const double tw = f[ii * 3 + jj] - coor[jj]; wR += tw * tw; }
            while (l >= 1) {
                const int idx1 = g[l], idx2 = g[l - 1];
                const double f1 = f[idx1 * 3 + AX];
                const double a = f1 - f[idx2 * 3 + AX], b = fd - f1, c = a + b;
                double uR = 0.0, vR = 0.0;
                for (int jj = 0; jj < 3; ++jj)
                    if (jj != AX) {
                        const double tu = f[idx2 * 3 + jj] - coor[jj], tv = f[idx1 * 3 + jj] - coor[jj];
                        uR += tu * tu;
                        vR += tv * tv;
                    }
                if (c * vR - b * uR - a * wR - a * b * c <= 0.0) break;
                --l;
            }
            ++l;
            g[l] = ii;
        }
        const int maxl = l;
        if (maxl >= 0) {
            // query walk: the nearest feature of every position of the line
            l = 0;
            for (int ii = 0; ii < len; ++ii) {
                double delta1 = 0.0;
                for (int jj = 0; jj < 3; ++jj) {
                    const double t = jj == AX ? f[g[l] * 3 + jj] - ii : f[g[l] * 3 + jj] - coor[jj];
                    delta1 += t * t;
                }
                while (l < maxl) {
                    double delta2 = 0.0;
                    for (int jj = 0; jj < 3; ++jj) {
                        const double t = jj == AX ? f[g[l + 1] * 3 + jj] - ii : f[g[l + 1] * 3 + jj] - coor[jj];
                        delta2 += t * t;
                    }
                    if (delta1 <= delta2) break;
                    delta1 = delta2;
                    ++l;
                }
                const int idx1 = g[l];
                const int64_t v = base + ii * strides[AX];
                F[v] = f[idx1 * 3]; F[n + v] = f[idx1 * 3 + 1]; F[2 * n + v] = f[idx1 * 3 + 2];
            }
        } else if (AX == 0) {
            // no voxel of the set on this line yet: scipy marks the first coordinate -1 and leaves the others at their
            // initial 0 (they stay that way if the whole volume is empty)
            for (int ii = 0; ii < len; ++ii) {
                const int64_t v = base + ii * strides[AX];
                F[v] = -1; F[n + v] = 0; F[2 * n + v] = 0;
            }
        }
    }
}

static unsigned mask_grid(int64_t n, int block) {
    int64_t b = ca_divup(n, block);
    if (b > 148 * 32) b = 148 * 32;
    return (unsigned)(b < 1 ? 1 : b);
}

extern "C" int ca_mask_erode(const float* mask, int H, int W, int D, float* out, ca_stream_t stream) {
    CA_CHECK_ARG(mask && out && H > 0 && W > 0 && D > 0, "mask_erode: bad arguments");
    CA_LAUNCH(k_mask_erode, dim3(mask_grid((int64_t)H * W * D, 256)), dim3(256), 0, (cudaStream_t)stream, mask, H, W, D, out);
    CA_CHECK_LAUNCH("mask_erode");
    return CA_OK;
}

extern "C" size_t ca_nearest_inside_workspace_bytes(int h, int w, int d) {
    if (h <= 0 || w <= 0 || d <= 0) return 0;
    return (size_t)h * w * d * 4 * sizeof(int);
}

// idx (3,h,w,d) int32 <- coordinates of the nearest voxel with inside != 0, scipy's choice among equidistant ones
// (distance_transform_edt(inside == 0, return_indices=True)[1]).  A volume without any inside voxel yields idx[0] = -1.
extern "C" int ca_nearest_inside(const float* inside, int h, int w, int d, int32_t* idx, void* workspace, size_t workspace_bytes,
                                 ca_stream_t stream) {
    CA_CHECK_ARG(inside && idx && workspace && h > 0 && w > 0 && d > 0, "nearest_inside: bad arguments");
    CA_CHECK_ARG((int64_t)h * w * d < (1ll << 29), "nearest_inside: volume too large");
    if (workspace_bytes < ca_nearest_inside_workspace_bytes(h, w, d)) {
        ca_set_error("nearest_inside: workspace too small");
        return CA_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int* scratch = (int*)workspace;
    CA_LAUNCH(k_ft_line<0>, dim3(mask_grid((int64_t)w * d, 128)), dim3(128), 0, st, inside, (int*)idx, h, w, d, scratch);
    CA_LAUNCH(k_ft_line<1>, dim3(mask_grid((int64_t)h * d, 128)), dim3(128), 0, st, inside, (int*)idx, h, w, d, scratch);
    CA_LAUNCH(k_ft_line<2>, dim3(mask_grid((int64_t)h * w, 128)), dim3(128), 0, st, inside, (int*)idx, h, w, d, scratch);
    CA_CHECK_LAUNCH("nearest_inside");
    return CA_OK;
}
