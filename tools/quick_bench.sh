#!/bin/bash
# quick GPU iteration: parity tests (subset) + bench without e2e/cpu legs
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('value %.3e  ms/step %.4f  frac %.3f  passes_us' % (d['value'], d['ms_per_step'], d['roofline']['frac']), {k: round(v*1000,1) for k,v in d['passes_ms'].items()})
    else:
        print(line)
"
