#!/bin/bash
# Variant build: libpapc_b200_<name>.so = the regular objects with the named sources recompiled under extra flags.
#   bash tools/build_variant.sh tri "-DPAPC_TT_TRIAGE" sa_mlp_tt.cu
#   PAPC_B200_LIB=papc_b200/lib/libpapc_b200_tri.so python tools/prof_layer.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2; shift 2
python papc_b200/csrc/build.py > /dev/null
od=papc_b200/lib/obj_var_$name; mkdir -p $od
objs=$(ls papc_b200/lib/obj/*.o)
for s in "$@"; do
  o=$od/${s%.cu}.o
  extra=""; [[ "$s" == fps.cu || "$s" == nms.cu ]] && [[ "$flags" != *fmad* ]] && extra="-fmad=false"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I papc_b200/csrc \
       $extra $flags -c papc_b200/csrc/$s -o $o &
  objs=$(echo "$objs" | grep -v "/${s%.cu}.o"); objs="$objs $o"
done
wait
nvcc -shared -o papc_b200/lib/libpapc_b200_$name.so $objs -lcudart
ls -la papc_b200/lib/libpapc_b200_$name.so
