#!/bin/bash
# quick regression + bench after a kernel change
timeout 600 python -m pytest tests/test_gpu_sa.py tests/test_gpu_models.py tests/test_gpu_fp.py tests/test_gpu_tc.py tests/test_gpu_fullsize.py tests/test_gpu_reference_models.py -m gpu -q -x --timeout 400 -p no:cacheprovider 2>&1 | tail -4
timeout 200 python bench.py --steps 30 --warmup 5 --no-extra 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step']); print([(k['name'][:34],k['M'],k['cin'],k['cout'],k['avg_ms']) for k in d['roofline']['kernels'] if k['name'].startswith('mlp')])"
