#!/bin/bash
# Triage build of the chained-MLP kernel: libpapc_b200_triage.so = the regular objects with sa_chain.cu recompiled
# under -DPAPC_CHAIN_TRIAGE (per-tile clock64 stamps, printed per launch when PAPC_CHAIN_CLK=1).
#   bash tools/build_triage.sh && PAPC_B200_LIB=papc_b200/lib/libpapc_b200_triage.so PAPC_CHAIN_CLK=1 python tools/chain_check.py ...
set -e
cd "$(dirname "$0")/.."
python papc_b200/csrc/build.py > /dev/null
mkdir -p papc_b200/lib/obj_triage
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I papc_b200/csrc \
     -DPAPC_CHAIN_TRIAGE -c papc_b200/csrc/sa_chain.cu -o papc_b200/lib/obj_triage/sa_chain.o
objs=$(ls papc_b200/lib/obj/*.o | grep -v sa_chain.o)
nvcc -shared -o papc_b200/lib/libpapc_b200_triage.so $objs papc_b200/lib/obj_triage/sa_chain.o -lcudart 2>/dev/null
ls -la papc_b200/lib/libpapc_b200_triage.so
