#!/bin/bash
# round 2, call 10 (1 GPU): deferred zero prefill (beside compositing only) + compositing launch bounds 5: tests and A/B
mkdir -p gpurun_out/r2c10 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c10
timeout 900 python -m pytest tests -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run base
run noprefill B2S_PREFILL=0
run c4 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_c4.so
run c1 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_c1.so
run base2
python - <<'PY'
import json
for f in ("base","noprefill","c4","c1","base2"):
    try:
        d=json.loads(open(f"gpurun_out/r2c10/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
