#!/bin/bash
# developer helper: build a variant of the library with extra -D flags for A/B runs on the GPU box
#   tools/build_variant.sh NAME -DKC_TILE_SLOTS=2 ...   ->  kompass-core_b200/lib/variants/libkompass_b200_NAME.so
# run with KOMPASS_B200_LIB=<that file>
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
OUT=kompass-core_b200/lib/variants; mkdir -p $OUT/obj_$NAME
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=default -cudart static"
for s in kc_api kc_planner kc_mapper kc_dwa; do
  nvcc $FLAGS "$@" -c kompass-core_b200/csrc/$s.cu -o $OUT/obj_$NAME/$s.o &
done
wait
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o $OUT/libkompass_b200_$NAME.so $OUT/obj_$NAME/*.o
rm -rf $OUT/obj_$NAME
echo $OUT/libkompass_b200_$NAME.so
