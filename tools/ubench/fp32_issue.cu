// Micro-benchmark: FP32 add / mul / fma issue rates on sm_100a, scalar vs packed f32x2 (add.rn.f32x2 etc.).
// Decides how the exact-order box filters of correlate.cu are written.  Build: see tools/ubench/run.sh
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITERS 4096
#define NCH 16

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
                 : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 r;
    asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc;}"
                 : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm volatile("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
                 : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float x, float y, long long* cyc) {
    float a[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) a[i] = threadIdx.x * 1e-3f + i;
    __shared__ float sh[1024];
    if (MODE == 6 || MODE == 7) { for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = x * i; __syncthreads(); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {            // scalar FADD
#pragma unroll
            for (int i = 0; i < NCH; ++i) a[i] = __fadd_rn(a[i], x);
        } else if (MODE == 1) {     // packed FADD2
#pragma unroll
            for (int i = 0; i < NCH; i += 2) { float2 r = add2(make_float2(a[i], a[i + 1]), make_float2(x, y)); a[i] = r.x; a[i + 1] = r.y; }
        } else if (MODE == 2) {     // scalar FFMA
#pragma unroll
            for (int i = 0; i < NCH; ++i) a[i] = __fmaf_rn(a[i], x, y);
        } else if (MODE == 3) {     // packed FFMA2
#pragma unroll
            for (int i = 0; i < NCH; i += 2) { float2 r = fma2(make_float2(a[i], a[i + 1]), make_float2(x, x), make_float2(y, y)); a[i] = r.x; a[i + 1] = r.y; }
        } else if (MODE == 4) {     // scalar sub, mul, add (raw SSD pattern): 3 ops per element
#pragma unroll
            for (int i = 0; i < NCH; ++i) { float d = __fsub_rn(x, a[i] * 0.f + y + i); a[i] = __fadd_rn(a[i], __fmul_rn(d, d)); }
        } else if (MODE == 5) {     // packed sub, mul, add
#pragma unroll
            for (int i = 0; i < NCH; i += 2) {
                float2 d = add2(make_float2(x, x), make_float2(-a[i], -a[i + 1]));
                float2 s = mul2(d, d);
                float2 r = add2(make_float2(a[i], a[i + 1]), s);
                a[i] = r.x; a[i + 1] = r.y;
            }
        } else if (MODE == 6) {     // scalar FADD with an LDS.128 per 12 adds (box-filter mix)
#pragma unroll
            for (int i = 0; i < NCH; i += 4) {
                float4 v = *reinterpret_cast<const float4*>(&sh[((threadIdx.x + it + i) & 255) * 4]);
                a[i] = __fadd_rn(__fadd_rn(__fadd_rn(a[i], v.x), v.y), v.z);
                a[i + 1] = __fadd_rn(__fadd_rn(__fadd_rn(a[i + 1], v.y), v.z), v.w);
                a[i + 2] = __fadd_rn(__fadd_rn(__fadd_rn(a[i + 2], v.z), v.w), v.x);
                a[i + 3] = __fadd_rn(__fadd_rn(__fadd_rn(a[i + 3], v.w), v.x), v.y);
            }
        } else if (MODE == 7) {     // packed FADD2 with an LDS.128 per 12 adds
#pragma unroll
            for (int i = 0; i < NCH; i += 4) {
                float4 v = *reinterpret_cast<const float4*>(&sh[((threadIdx.x + it + i) & 255) * 4]);
                float2 p = make_float2(a[i], a[i + 1]), q = make_float2(a[i + 2], a[i + 3]);
                p = add2(add2(add2(p, make_float2(v.x, v.y)), make_float2(v.y, v.z)), make_float2(v.z, v.w));
                q = add2(add2(add2(q, make_float2(v.z, v.w)), make_float2(v.w, v.x)), make_float2(v.x, v.y));
                a[i] = p.x; a[i + 1] = p.y; a[i + 2] = q.x; a[i + 3] = q.y;
            }
        } else if (MODE == 8) {     // half the chains FADD, half FMUL (do they share a pipe?)
#pragma unroll
            for (int i = 0; i < NCH; i += 2) { a[i] = __fadd_rn(a[i], x); a[i + 1] = __fmul_rn(a[i + 1], y); }
        } else if (MODE == 9) {     // scalar FMUL
#pragma unroll
            for (int i = 0; i < NCH; ++i) a[i] = __fmul_rn(a[i], y);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread, int blocks_per_sm) {
    float* out; long long* cyc;
    const int nb = 148 * blocks_per_sm;
    cudaMalloc(&out, (size_t)nb * 256 * 4);
    cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<nb, 256>>>(out, 1.0001f, 0.9999f, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<nb, 256>>>(out, 1.0001f, 0.9999f, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops = ops_per_iter_per_thread * ITERS * 256.0 * blocks_per_sm;   // per SM
    printf("%-44s blocks/SM=%d  %8.3f ms  block0 cycles=%lld  -> %7.1f fp32 ops/clk/SM (block-0 clock), %6.2f Tops/s chip\n", name,
           blocks_per_sm, ms, c, ops / (double)c, ops * 148 / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int b : {4, 8}) {
        if (b == 4) {
            run<0>("FADD scalar", NCH, 4);  run<1>("FADD2 packed", NCH, 4);  run<9>("FMUL scalar", NCH, 4);
            run<2>("FFMA scalar (1 op each)", NCH, 4);  run<3>("FFMA2 packed (1 op each)", NCH, 4);
            run<8>("FADD+FMUL alternating", NCH, 4);
            run<4>("sub,mul,add scalar (+2 aux)", NCH * 3, 4);  run<5>("sub,mul,add packed", NCH * 3, 4);
            run<6>("box mix scalar: 12 FADD per LDS.128", 12 * NCH / 4, 4);  run<7>("box mix packed: 6 FADD2 per LDS.128", 12 * NCH / 4, 4);
        } else {
            run<0>("FADD scalar", NCH, 8);  run<1>("FADD2 packed", NCH, 8);
            run<6>("box mix scalar: 12 FADD per LDS.128", 12 * NCH / 4, 8);  run<7>("box mix packed: 6 FADD2 per LDS.128", 12 * NCH / 4, 8);
        }
    }
    return 0;
}
