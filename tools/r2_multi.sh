#!/bin/bash
# multi-GPU call (gpurun --gpus N): protocol check + transports.  usage: tools/r2_multi.sh N "transports" [steps]
N=${1:-2}; GATHERS=${2:-"ce mc none"}; STEPS=${3:-40}; O=gpurun_out/r2m$N; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
echo "== multi_gpu_check world $N"; timeout 600 bash -c "$(declare -f run); N=$N; run tests/multi_gpu_check.py" 2>&1 | grep -E "MULTI_GPU_CHECK|Error|error|assert" | head -6 | tee $O/multi_gpu_check.txt
for g in $GATHERS; do
  timeout 600 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps $STEPS --warmup 3 --no-e2e --no-cpu-baseline --no-extras --gather $g" 2>$O/err_$g.log | tail -1 > $O/bench_$g.json
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$g.json").read().strip().splitlines()[-1])
    t=d["timing"]
    print("$g", "value %.3e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v*1000,1) for k,v in d["passes_ms"].items()}, "graph", t.get("graph_ms_per_step"), "eager", round(t["eager_ms_per_step"],4), "plain", round(t["eager_no_events_ms_per_step"],4), t.get("graph_unavailable"), d.get("gather_transport"), d.get("gather_note"), "verified", d.get("multi_gpu_verified"), d.get("multi_gpu_verified_error"))
except Exception as e:
    print("$g", "FAILED", e); import subprocess; print(subprocess.run("tail -5 $O/err_$g.log", shell=True, capture_output=True, text=True).stdout)
PY
done
ls $O
