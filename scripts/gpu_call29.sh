#!/bin/bash
# round 2, call 29 (1 GPU): bwd_classify with 128-bit accesses and 4096 rows per CTA (A/B against the scalar kernel)
mkdir -p gpurun_out/r2c29 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c29
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run base
run scalar B2S_CLASSIFY_SCALAR=1
run base2
python - <<'PY'
import json
for f in ("base","scalar","base2"):
    try:
        d=json.loads(open(f"gpurun_out/r2c29/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
