#!/bin/bash
# activation-image mode check: parity tests that cover small-M fp16 layers, then the bench A/B
timeout 400 python -m pytest tests/test_gpu_sa.py tests/test_gpu_models.py tests/test_gpu_chain.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
timeout 400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_fp.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -3
for f in 1 0; do PAPC_TT_XIMG=$f timeout 200 python bench.py --steps 30 --warmup 5 --no-extra 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ximg=$f', d['value'], d['ms_per_step']); print([(k['name'][:34],k['M'],k['cin'],k['cout'],k['avg_ms']) for k in d['roofline']['kernels'] if k['M']==4096 or k['name'].startswith('pool')])"; done
