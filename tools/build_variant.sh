#!/bin/bash
# Kernel A/B experiments: rebuilds ONE translation unit with extra -D flags and links it with the objects of the
# regular build into phoenix_drone_simulation_b200/libphoenix_b200_<name>.so (select it with PDX_LIB=<path>).
#   tools/build_variant.sh <name> <unit.cu> [-DFOO=1 ...]
set -e
name=$1; unit=$2; shift 2
here=$(cd "$(dirname "$0")/.." && pwd)
pkg=$here/phoenix_drone_simulation_b200
extra=""
case $unit in *f32*|pdx_collect.cu) extra="-use_fast_math";; *f64*) extra="-fmad=false";; esac
obj=$pkg/build/variant_${name}_${unit%.cu}.o
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC $extra "$@" -c $pkg/csrc/$unit -o $obj
objs=""
for o in $pkg/build/pdx_*.o; do b=$(basename $o); [ "$b" = "${unit%.cu}.o" ] || [ "$b" = "pdx_policy_tc_timing.o" ] || objs="$objs $o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $pkg/libphoenix_b200_$name.so $objs $obj
echo $pkg/libphoenix_b200_$name.so
