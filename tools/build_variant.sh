#!/bin/bash
# tuning aid: build a variant of the library from csrc/ with one file replaced:  tools/build_variant.sh <name> <file.cu> <replacement>
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/convexadam_b200/csrc
OUT=$ROOT/convexadam_b200/variants
mkdir -p $OUT
TMP=$(mktemp -d)
cp $SRC/*.cu $SRC/*.h $SRC/*.cuh $TMP/
mkdir -p $TMP/../../include 2>/dev/null
cp "$3" $TMP/$2
sed -i "s|#include \"../../include/convexadam_b200.h\"|#include \"$ROOT/include/convexadam_b200.h\"|" $TMP/ca_platform.h
(cd $TMP && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=true -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared -cudart static common.cu pool.cu mind.cu correlate.cu coupled.cu field.cu adam.cu -o $OUT/lib_$1.so -ccbin /usr/bin/g++ 2>&1 | grep -iE "error" )
ls -la $OUT/lib_$1.so | awk '{print $NF, $5}'
