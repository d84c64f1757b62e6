// Probe: which tensor-map geometries does cp.async.bulk.tensor accept on this GPU?  (debug aid for ca_tma.cuh)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct alignas(64) TM { CUtensorMap m; };

__global__ void k4(const __grid_constant__ TM tm, int c0, int c1, int c2, int c3, int nfloats, float* out) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ unsigned long long bar;
    float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~(uintptr_t)127);
    unsigned b = (unsigned)__cvta_generic_to_shared(&bar), d = (unsigned)__cvta_generic_to_shared(sm);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(nfloats * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(d),
                     "l"(&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(b) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = sm[i];
}

int main() {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no entry point\n"); return 1; }
    EncodeFn enc = (EncodeFn)fn;
    const int d = 40, w = 21, h = 7, C = 12;
    std::vector<float> host((size_t)d * w * h * C);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (float)(i % 1000) + 1.f;
    float *dev, *out;
    cudaMalloc(&dev, host.size() * 4); cudaMalloc(&out, 1 << 20);
    cudaMemcpy(dev, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    struct Case { int bz, by, bc, c0, c1, c2, c3; CUtensorMapL2promotion l2; const char* name; };
    const Case cases[] = {
        {32, 8, 2, 0, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE, "inner 32, in bounds"},
        {36, 8, 2, 0, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE, "inner 36, in bounds"},
        {36, 20, 6, 0, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE, "inner 36 x 20 x 6, in bounds"},
        {36, 20, 6, -4, -2, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE, "inner 36 x 20 x 6, negative start -4"},
        {36, 20, 6, -8, -2, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "start -8, L2 promotion 256B"},
        {36, 20, 6, 28, 14, 5, 6, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "far corner"},
        {40, 20, 6, -4, -2, -2, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "inner 40, x negative"},
    };
    for (const Case& c : cases) {
        TM tm;
        const cuuint64_t dims[4] = {d, w, h, C};
        const cuuint64_t strides[3] = {(cuuint64_t)d * 4, (cuuint64_t)d * w * 4, (cuuint64_t)d * w * h * 4};
        const cuuint32_t box[4] = {(cuuint32_t)c.bz, (cuuint32_t)c.by, 1u, (cuuint32_t)c.bc};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm.m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dev, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, c.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int nf = c.bz * c.by * c.bc;
        cudaMemset(out, 0, nf * 4);
        k4<<<1, 128, nf * 4 + 256>>>(tm, c.c0, c.c1, c.c2, c.c3, nf, out);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> got(nf);
        cudaMemcpy(got.data(), out, nf * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int cc = 0; cc < c.bc; ++cc) for (int y = 0; y < c.by; ++y) for (int z = 0; z < c.bz; ++z) {
            const int iz = c.c0 + z, iy = c.c1 + y, ix = c.c2, ic = c.c3 + cc;
            const bool in = iz >= 0 && iz < d && iy >= 0 && iy < w && ix >= 0 && ix < h && ic >= 0 && ic < C;
            const float want = in ? host[(((size_t)ic * h + ix) * w + iy) * d + iz] : 0.f;
            if (got[((size_t)cc * c.by + y) * c.bz + z] != want) ++bad;
        }
        printf("%-40s encode=%d run=%s mismatches=%d\n", c.name, (int)r, cudaGetErrorString(e), bad);
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 2; }
    }
    return 0;
}
// result on B200 (driver 580): every geometry above loads correctly as long as the innermost start coordinate is a multiple
// of 4 floats; a start of -2 (8 bytes) raises "an illegal instruction was encountered".
