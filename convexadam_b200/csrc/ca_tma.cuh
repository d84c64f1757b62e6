// Tensor-map TMA for 4-D feature tiles: cp.async.bulk.tensor.4d (UTMALDG) with out-of-bounds zero fill.
//
// The reference pads the moving features with zeros (F.pad at convex_adam_utils.py:76) and unfolds them; here a tile of
// the UNPADDED (C,h,w,d) volume is requested at signed coordinates and the TMA unit fills whatever lies outside the
// tensor with zeros: no padded copy, no halo stores, no bounds checks in the kernel.  One request brings a box of
// (BZ, BY, 1, BC) elements = BC channels x BY rows x BZ voxels of one H plane.
//
// CA_EMU: the same call is a strided copy with zero fill, executed by the issuing fiber.
#pragma once
#include "ca_platform.h"

#ifdef CA_EMU
struct CaTensorMap {
    const float* base;
    int dim[4];            // elements, innermost first: (z, y, x, c)
    long long stride[4];   // elements
    int box[4];
};
static inline int ca_tmap_make(CaTensorMap* tm, const float* base, int d, int w, int h, int C, long long pitch_z, int bz, int by,
                               int bc) {
    tm->base = base;
    tm->dim[0] = d; tm->dim[1] = w; tm->dim[2] = h; tm->dim[3] = C;
    tm->stride[0] = 1; tm->stride[1] = pitch_z; tm->stride[2] = pitch_z * w; tm->stride[3] = pitch_z * w * h;
    tm->box[0] = bz; tm->box[1] = by; tm->box[2] = 1; tm->box[3] = bc;
    return CA_OK;
}
static inline void ca_tma_load_4d(float* dst, const CaTensorMap* tm, int c0, int c1, int c2, int c3, unsigned long long* bar) {
    (void)bar;
    if (c0 & 3) {   // the hardware raises "illegal instruction" unless the box starts on a 16-byte boundary of the row
        fprintf(stderr, "ca_tma_load_4d: innermost coordinate %d is not a multiple of 4 floats\n", c0);
        abort();
    }
    for (int c = 0; c < tm->box[3]; ++c)
        for (int x = 0; x < tm->box[2]; ++x)
            for (int y = 0; y < tm->box[1]; ++y)
                for (int z = 0; z < tm->box[0]; ++z) {
                    const int iz = c0 + z, iy = c1 + y, ix = c2 + x, ic = c3 + c;
                    const bool in = iz >= 0 && iz < tm->dim[0] && iy >= 0 && iy < tm->dim[1] && ix >= 0 && ix < tm->dim[2] && ic >= 0 && ic < tm->dim[3];
                    *dst++ = in ? tm->base[iz * tm->stride[0] + iy * tm->stride[1] + ix * tm->stride[2] + ic * tm->stride[3]] : 0.f;
                }
}
#define CA_TMAP_PARAM const CaTensorMap
#else
#include <cuda.h>
struct alignas(64) CaTensorMap {
    CUtensorMap m;
};
#define CA_TMAP_PARAM const __grid_constant__ CaTensorMap

// (C,h,w,d) float32 volume with row pitch `pitch_z` elements (a multiple of 4: tensor-map strides are multiples of 16
// bytes), base 16-byte aligned; box = (bz, by, 1, bc), bz a multiple of 4.
static inline int ca_tmap_make(CaTensorMap* tm, const float* base, int d, int w, int h, int C, long long pitch_z, int bz, int by,
                               int bc) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess) {
            ca_set_error("tensor map: cuTensorMapEncodeTiled not available from the driver");
            return CA_ERR_LAUNCH;
        }
        encode = (EncodeFn)fn;
    }
    if ((pitch_z & 3) || (((uintptr_t)base) & 15) || (bz & 3) || bz > 256 || by > 256 || bc > 256) {
        ca_set_error("tensor map: unsupported geometry (pitch %lld, box %d x %d x %d)", pitch_z, bz, by, bc);
        return CA_ERR_UNSUPPORTED;
    }
    const cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)C};
    const cuuint64_t strides[3] = {(cuuint64_t)pitch_z * 4, (cuuint64_t)pitch_z * w * 4, (cuuint64_t)pitch_z * w * h * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bz, (cuuint32_t)by, 1u, (cuuint32_t)bc};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = encode(&tm->m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);      // "NONE" = zeros (the other choice is NaN)
    if (r != CUDA_SUCCESS) {
        ca_set_error("tensor map: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return CA_ERR_LAUNCH;
    }
    return CA_OK;
}
// dst: 128-byte aligned shared memory; c0 (innermost coordinate, may be negative) a multiple of 4 floats: the box must start on
// a 16-byte boundary of the global row, anything else is an "illegal instruction" at run time (tools/ubench/tma_probe.cu);
// completes box-bytes on `bar`
static __device__ __forceinline__ void ca_tma_load_4d(float* dst, const CaTensorMap* tm, int c0, int c1, int c2, int c3,
                                                      unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
                     ca_smem_u32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(ca_smem_u32(bar))
                 : "memory");
}
static __device__ __forceinline__ void ca_tmap_prefetch(const CaTensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
#endif
