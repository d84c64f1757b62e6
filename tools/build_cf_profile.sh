#!/bin/bash
# debug aid: variant of the library with -DCF_PROFILE (per-phase cycle counters in the fused correlate kernel)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/convexadam_b200/csrc
OUT=$ROOT/convexadam_b200/variants
mkdir -p $OUT
(cd $SRC && nvcc -DCF_PROFILE -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared -cudart static common.cu pool.cu mind.cu correlate.cu correlate_fused.cu coupled.cu field.cu adam.cu -o $OUT/lib_cfprof.so -ccbin /usr/bin/g++ 2>&1 | grep -iE "error|spill" )
ls -la $OUT/lib_cfprof.so | awk '{print $NF, $5}'
