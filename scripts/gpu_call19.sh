#!/bin/bash
# round 2, call 19 (1 GPU): counts copied to the host on a side stream before the depth sort is enqueued (the host no longer
# waits for the sort), histogram kernel 4 rows per iteration, own zero fill for the gradient slab -- full GPU suite, A/B
mkdir -p gpurun_out/r2c19 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c19
timeout 900 python -m pytest tests -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run fused
run unfused B2S_BINNING_FUSED=0
run fused2
timeout 300 python scripts/host_overhead.py > $O/host_overhead.txt 2>&1; echo "host_overhead rc=$?"; head -3 $O/host_overhead.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("fused","unfused","fused2"):
    try:
        d=json.loads(open(f"gpurun_out/r2c19/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
