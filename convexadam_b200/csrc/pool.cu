// avg_pool3d (count_include_pad=True) with torch-CPU summation order, and the nnUNet weighted one-hot pooling.
//
// Replaces F.avg_pool3d at convex_adam_MIND.py:118-119,149-150 (feature pooling), :166 (B-spline smoothing
// of the Adam parameters, via adam.cu), :191 (selected_smooth) and the weighted one-hot features of
// convex_adam_nnUNet.py:19-38.  HBM-bound elementwise/stencil work: one thread per output, D fastest so that
// a warp reads/writes contiguous rows.
#include "ca_platform.h"

__global__ void __launch_bounds__(256) k_avgpool3d(const float* __restrict__ in, int64_t nc, int iT, int iH, int iW,
                                                   int k, int stride, int pad, int oT, int oH, int oW,
                                                   float* __restrict__ out) {
    const int64_t total = nc * oT * oH * (int64_t)oW;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(o % oW);
        const int i = (int)((o / oW) % oH);
        const int t = (int)((o / ((int64_t)oW * oH)) % oT);
        const int64_t c = o / ((int64_t)oW * oH * oT);
        int ts = t * stride - pad, hs = i * stride - pad, ws = j * stride - pad;
        int te = ca_min(ts + k, iT + pad), he = ca_min(hs + k, iH + pad), we = ca_min(ws + k, iW + pad);
        const int pool = (te - ts) * (he - hs) * (we - ws);
        ts = ca_max(ts, 0); hs = ca_max(hs, 0); ws = ca_max(ws, 0);
        te = ca_min(te, iT); he = ca_min(he, iH); we = ca_min(we, iW);
        const float* ip = in + c * (int64_t)iT * iH * iW;
        float sum = 0.f;
        if (ts < te && hs < he && ws < we) {
            for (int z = ts; z < te; ++z)
                for (int y = hs; y < he; ++y)
                    for (int x = ws; x < we; ++x) sum = __fadd_rn(sum, __ldg(ip + ((int64_t)z * iH + y) * iW + x));
            sum = ca_div(sum, (float)pool);
        }
        out[o] = sum;
    }
}

static inline int grid_for(int64_t total, int block, int max_blocks = 148 * 16) {
    int64_t b = ca_divup(total, block);
    if (b > max_blocks) b = max_blocks;
    if (b < 1) b = 1;
    return (int)b;
}

int ca_launch_avgpool3d(const float* in, int64_t nc, int T, int H, int W, int k, int stride, int pad, float* out,
                        cudaStream_t st) {
    const int oT = (T + 2 * pad - k) / stride + 1, oH = (H + 2 * pad - k) / stride + 1, oW = (W + 2 * pad - k) / stride + 1;
    const int64_t total = nc * oT * oH * (int64_t)oW;
    if (total <= 0) return CA_OK;
    CA_LAUNCH(k_avgpool3d, dim3(grid_for(total, 256)), dim3(256), 0, st, in, nc, T, H, W, k, stride, pad, oT, oH, oW, out);
    CA_CHECK_LAUNCH("avgpool3d");
    return CA_OK;
}

extern "C" int ca_avgpool3d(const float* in, int64_t nc, int T, int H, int W, int k, int stride, int pad, float* out,
                            ca_stream_t stream) {
    CA_CHECK_ARG(in && out, "avgpool3d: null pointer");
    CA_CHECK_ARG(nc > 0 && T > 0 && H > 0 && W > 0 && k > 0 && stride > 0 && pad >= 0, "avgpool3d: bad dims");
    CA_CHECK_ARG(pad * 2 <= k, "avgpool3d: pad should be at most half of kernel size");
    CA_CHECK_ARG(T + 2 * pad >= k && H + 2 * pad >= k && W + 2 * pad >= k, "avgpool3d: kernel larger than padded input");
    return ca_launch_avgpool3d(in, nc, T, H, W, k, stride, pad, out, (cudaStream_t)stream);
}

// nnUNet features: out[c, x] = mean over the gs^3 block of (lut[label]==c ? weight[c] : 0), sequential block sum / gs^3
__global__ void __launch_bounds__(256) k_onehot_pooled(const float* __restrict__ labels, int H, int W, int D,
                                                       const int32_t* __restrict__ lut, int n_lut,
                                                       const float* __restrict__ weight, int C, int gs, int h, int w,
                                                       int d, float* __restrict__ out) {
    const int64_t n = (int64_t)h * w * d;
    const int64_t total = n * C;
    const float cnt = (float)(gs * gs * gs);
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = o % n;
        const int c = (int)(o / n);
        const int z = (int)(v % d), y = (int)((v / d) % w), x = (int)(v / ((int64_t)d * w));
        const float wc = weight[c];
        float sum = 0.f;
        for (int a = 0; a < gs; ++a)
            for (int b = 0; b < gs; ++b)
                for (int e = 0; e < gs; ++e) {
                    const float lf = __ldg(labels + ((int64_t)(x * gs + a) * W + (y * gs + b)) * D + (z * gs + e));
                    const int li = (int)lf;
                    const int ch = (li >= 0 && li < n_lut) ? lut[li] : -1;
                    sum = __fadd_rn(sum, ch == c ? wc : 0.f);
                }
        out[o] = ca_div(sum, cnt);
    }
}

extern "C" int ca_onehot_pooled(const float* labels, int H, int W, int D, const int32_t* lut, int n_lut,
                                const float* weight, int C, int gs, float* out, ca_stream_t stream) {
    CA_CHECK_ARG(labels && lut && weight && out, "onehot_pooled: null pointer");
    CA_CHECK_ARG(H > 0 && W > 0 && D > 0 && C > 0 && gs > 0 && n_lut > 0, "onehot_pooled: bad dims");
    const int h = H / gs, w = W / gs, d = D / gs;
    CA_CHECK_ARG(h > 0 && w > 0 && d > 0, "onehot_pooled: volume smaller than the pooling block");
    const int64_t total = (int64_t)h * w * d * C;
    CA_LAUNCH(k_onehot_pooled, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, labels, H, W, D, lut, n_lut,
              weight, C, gs, h, w, d, out);
    CA_CHECK_LAUNCH("onehot_pooled");
    return CA_OK;
}
