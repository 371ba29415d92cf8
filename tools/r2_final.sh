#!/bin/bash
# round 2, final single-GPU call: parity suite, the default bench line, C2 / C3 lines, ncu launch list + full capture of one
# frame, a short compute-sanitizer pass.  Results under gpurun_out/r2z/ (copied to profiles/ by hand).
O=gpurun_out/r2z; mkdir -p $O
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest_gpu.txt
python bench.py --steps 50 --warmup 3 2>>$O/err.log | tail -1 > $O/bench_full.json
python - <<PY
import json
d=json.loads(open("$O/bench_full.json").read().strip().splitlines()[-1])
print("full", "value %.3e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v*1000,1) for k,v in d["passes_ms"].items()}, "roofline", round(d["roofline"]["frac"],3), "e2e", d["e2e"]["ms_per_step"], "inc", d["e2e_incremental"]["ms_per_step"])
PY
for w in C2 C3; do python bench.py --workload $w --steps 30 --warmup 3 --no-e2e --no-cpu-baseline --no-extras 2>>$O/err.log | tail -1 > $O/bench_$w.json; python -c "
import json; d=json.loads(open('$O/bench_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['ms_per_step'],4), {k: round(v*1000,1) for k,v in d['passes_ms'].items()}, d.get('raster_ms'))"; done
python bench.py --impl reference --steps 3 --warmup 1 2>>$O/err.log | tail -1 > $O/bench_reference.json; cut -c1-300 $O/bench_reference.json
echo "== ncu full: every kernel of one frame"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"drawcull_kernel|clustercull_filter_kernel|pyramid_kernel|footprint_kernel" -s 18 -c 6 -f -o $O/prof_frame python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 24 --csv --log-file $O/launches.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
echo "== compute-sanitizer memcheck (subset)"
SEL="kitten_4096 or tiny or overflow or taskcull or hostile or big_meshes or decode or produced_depth"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error:" | head -6 | tee $O/sanitizer.txt
ls -la $O; tail -3 $O/err.log
