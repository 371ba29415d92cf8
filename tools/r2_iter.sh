#!/bin/bash
# round 2, iteration call: GPU parity suite, a quick C4 bench, optional A/B variants, optional ncu capture of chosen kernels.
# usage: tools/r2_iter.sh TAG [ncu-kernel-regex] [variant ...]      results under gpurun_out/TAG/
TAG=${1:-iter}; KREGEX=${2:-}; shift; shift
O=gpurun_out/$TAG; mkdir -p $O
q() { python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>>$O/err.log | tail -1; }
summ() { python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'value %.3e ms/step %.4f' % (d['value'], d['ms_per_step']), {k: round(v*1000,1) for k,v in d['passes_ms'].items()}, (d.get('cluster_filter') or {}).get('exact_share'), d['timing'].get('eager_no_events_ms_per_step'))
"; }
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_gpu.txt
q > $O/bench_default.json; summ default < $O/bench_default.json
for v in "$@"; do NVC_LIB_PATH=$PWD/niagara_b200/variant_$v.so q > $O/bench_$v.json; summ $v < $O/bench_$v.json; done
python bench.py --workload C2 --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>>$O/err.log | tail -1 > $O/bench_c2.json; summ C2 < $O/bench_c2.json
if [ -n "$KREGEX" ]; then
  echo "== ncu full: $KREGEX"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s 12 -c 4 -f -o $O/prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log | cut -c1-200
fi
tail -5 $O/err.log
