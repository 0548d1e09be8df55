export PAPC_B200_LIB=papc_b200/lib/libpapc_b200_tri.so
for d in 0 16 1 2 4 3 6 7; do
  echo "== PAPC_TT_DBG=$d"
  PAPC_TT_DBG=$d timeout 200 python tools/prof_layer.py sa1.l2 sa1.l3 sa2.l1 sa2.l2 sa2.l3 2>&1 | tail -6
done
