#!/bin/bash
# debug aid: variant of the library with -DCP_PROFILE (per-phase cycle counters in the producer / consumer correlate kernel)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/convexadam_b200/csrc
OUT=$ROOT/convexadam_b200/variants
mkdir -p $OUT
(cd $SRC && env -u CC -u CXX nvcc -DCP_PROFILE $CP_EXTRA -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared -cudart static common.cu pool.cu mind.cu correlate.cu correlate_fused.cu correlate_pipe.cu coupled.cu field.cu adam.cu mask.cu -o $OUT/lib_cpprof.so -ccbin /usr/bin/g++ 2>&1 | grep -iE "error|spill" )
ls -la $OUT/lib_cpprof.so | awk '{print $NF, $5}'
