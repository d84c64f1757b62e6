/* Exhaustive proof that ca_div27 (convexadam_b200/csrc/ca_platform.h) equals the IEEE fp32 division x/27 for every
 * non-negative finite float, subnormals included (the operations are sign-symmetric).
 *   gcc -O2 -mfma -ffp-contract=off -fopenmp tests/div27_exhaustive.c -o /tmp/div27 -lm && /tmp/div27 [stride]
 * stride 1 (default) walks all 2^31 bit patterns (~3 min on 8 cores); tests/test_host_logic.py runs stride 4099.
 * Result recorded in DESIGN.md: 27, 8, 64, 125, 343 -> 0 mismatches; 12 and 216 mismatch only below 1e-37. */
#include <math.h>
#include <stdlib.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <omp.h>
// exhaustive check: q1 = fma(r, c, q0), r = fma(-27, q0, x), q0 = x*c, c = RN(1/27)  ==  x / 27 (IEEE RN) ?
static inline float div27(float x, float c, float dv) { float q0 = x * c; float r = fmaf(-dv, q0, x); return fmaf(r, c, q0); }
int main(int argc, char** argv) {
    const uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
    int rc = 0;
    const float dvs[] = {27.f, 8.f, 216.f, 125.f, 64.f, 12.f, 343.f};
    for (int k = 0; k < (stride > 1 ? 1 : 7); ++k) {
        const float dv = dvs[k], c = 1.0f / dv;
        uint64_t bad = 0; uint32_t firstbad = 0, lastbad = 0;
#pragma omp parallel for reduction(+:bad) schedule(static)
        for (uint64_t u = 0; u < 0x7f800000ull; u += stride) {
            uint32_t ui = (uint32_t)u; float x; memcpy(&x, &ui, 4);
            float a = x / dv, b = div27(x, c, dv);
            uint32_t ai, bi; memcpy(&ai, &a, 4); memcpy(&bi, &b, 4);
            if (ai != bi) { bad++; 
#pragma omp critical
                { if (!firstbad || ui < firstbad) firstbad = ui; if (ui > lastbad) lastbad = ui; } }
        }
        float fb, lb; memcpy(&fb, &firstbad, 4); memcpy(&lb, &lastbad, 4);
        if (k == 0 && bad) rc = 1;
        printf("div %g: mismatches %llu  first %g (0x%08x) last %g (0x%08x)\n", dv, (unsigned long long)bad, fb, firstbad, lb, lastbad);
    }
    return rc;
}
