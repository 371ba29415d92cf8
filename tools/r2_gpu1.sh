#!/bin/bash
# round 2, first GPU call: parity of the filtered cluster kernel on hardware, A/B against the exact kernel, occupancy variants,
# the round-1 build flags (PDL, SMEM_ITEMS), pending GPU tests, ncu of the new kernels.  Results under gpurun_out/r2a/.
O=gpurun_out/r2a; mkdir -p $O
q() { python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu-baseline $X 2>>$O/err.log | tail -1; }
summ() { python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'value %.3e ms/step %.4f' % (d['value'], d['ms_per_step']), {k: round(v*1000,1) for k,v in d['passes_ms'].items()}, d.get('cluster_filter'))
"; }
echo "== pytest -m gpu (filtered kernel default)"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_gpu.txt
echo "== pending gpu tests"; timeout 600 python -m pytest tests/pending/gpu_end_to_end.py -q -o python_functions='pending_test_*' 2>&1 | tail -4 | tee $O/pytest_pending.txt
X= q > $O/bench_filter.json; summ filter < $O/bench_filter.json; X=--no-extras
NVC_PREPARE_HIZ=0 q > $O/bench_filter_nofp.json; summ filter_nofp < $O/bench_filter_nofp.json
NVC_DRAW_FILTER=0 q > $O/bench_nodrawfilter.json; summ no_draw_filter < $O/bench_nodrawfilter.json
NVC_CLUSTER_FILTER=0 q > $O/bench_exact.json; summ exact < $O/bench_exact.json
for v in fb3 fb5 fb6 pdl dpt1 dpt4; do NVC_LIB_PATH=$PWD/niagara_b200/variant_$v.so q > $O/bench_$v.json; summ $v < $O/bench_$v.json; done
NVC_CLUSTER_FILTER=0 NVC_LIB_PATH=$PWD/niagara_b200/variant_smem_items.so q > $O/bench_exact_smem_items.json; summ exact_smem_items < $O/bench_exact_smem_items.json
NVC_CLUSTER_FILTER=0 NVC_LIB_PATH=$PWD/niagara_b200/variant_pdl.so q > $O/bench_exact_pdl.json; summ exact_pdl < $O/bench_exact_pdl.json
echo "== pytest -m gpu with the PDL build"; NVC_LIB_PATH=$PWD/niagara_b200/variant_pdl.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | tee $O/pytest_gpu_pdl.txt
python bench.py --workload C2 --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>>$O/err.log | tail -1 > $O/bench_c2.json; summ C2 < $O/bench_c2.json
echo "== ncu full, new kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"clustercull_filter_kernel|footprint_kernel" -s 12 -c 3 -f -o $O/prof_filter python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 20 --csv --log-file $O/launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
ls -la $O
python bench.py --steps 50 --warmup 3 2>>$O/err.log | tail -1 > $O/bench_full.json; summ full < $O/bench_full.json
python bench.py --workload C3 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>>$O/err.log | tail -1 > $O/bench_c3.json; summ C3 < $O/bench_c3.json
echo "== compute-sanitizer (filtered kernels, drawcull DPT, raster, footprint)"
SEL="kitten_4096 or tiny or pyramid_sizes or overflow or taskcull or hostile or big_meshes or decode or without_prepared or produced_depth"
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|Error|error:|hazard" | head -8 | sed "s/^/$tool: /" | tee -a $O/sanitizer.txt
done
