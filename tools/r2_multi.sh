#!/bin/bash
# multi-GPU call (gpurun --gpus N): protocol + transports.  usage: tools/r2_multi.sh N
N=${1:-2}; O=gpurun_out/r2m$N; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
echo "== multi_gpu_check world $N"; timeout 900 bash -c "$(declare -f run); N=$N; run tests/multi_gpu_check.py" 2>&1 | tail -6 | tee $O/multi_gpu_check.txt
for g in ce mc fused nccl none; do
  timeout 600 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps 60 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --gather $g" 2>$O/err_$g.log | tail -1 > $O/bench_$g.json
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$g.json").read().strip().splitlines()[-1])
    print("$g", "value %.3e ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v*1000,1) for k,v in d["passes_ms"].items()}, d.get("gather_transport"), d.get("gather_note"), "verified", d.get("multi_gpu_verified"))
except Exception as e:
    print("$g", "FAILED", e)
PY
done
NCCL_DEBUG=INFO timeout 300 bash -c "$(declare -f run); N=$N; run bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-extras --gather nccl" 2>&1 | grep -i -m5 "nvls\|multicast" > $O/nccl_nvls.txt; cat $O/nccl_nvls.txt
ls $O
