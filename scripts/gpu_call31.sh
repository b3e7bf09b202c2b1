#!/bin/bash
# round 2, call 31 (1 GPU, the last GPU seconds of the round): bwd_classify with every load issued before the radius test (A/B)
mkdir -p gpurun_out/r2c31 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c31
B="python bench.py --gpus 1 --steps 40 --warmup 8 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 120 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run base
run dependent B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_clsdep.so
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sparse_backward or forward_backward or leak or keep_steps" > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python - <<'PY'
import json
for f in ("base","dependent"):
    try:
        d=json.loads(open(f"gpurun_out/r2c31/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
