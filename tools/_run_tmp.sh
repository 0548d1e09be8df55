timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_sa.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
echo "== tf32"; python tools/tt_triage.py 0,16,5,6,22 8,28
echo "== f16"; PAPC_TT_PREC=f16 python tools/tt_triage.py 0,16,5,6,22 8,28
python tools/prof_layer.py 2>&1 | tail -9
