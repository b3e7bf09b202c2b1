#!/bin/bash
# round 2, call 9 (1 GPU): side-stream zero prefill (tests + A/B), launch-bound variants of the sort and compositing kernels
mkdir -p gpurun_out/r2c9 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c9
timeout 900 python -m pytest tests -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run base
run noprefill B2S_PREFILL=0
run r4 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_r4.so
run c6 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_c6.so
run c5 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_c5.so
B="python bench.py --gpus 1 --steps 30 --warmup 10 --no-extras --no-e2e --no-cpu-baseline --workload config3_init"
run init_base
run init_c6 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_c6.so
python - <<'PY'
import json
for f in ("base","noprefill","r4","c6","c5","init_base","init_c6"):
    try:
        d=json.loads(open(f"gpurun_out/r2c9/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
