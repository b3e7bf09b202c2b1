#!/bin/bash
# round 2, call 23 (1 GPU): projection VJP of the live rows on a second stream beside the v_sh zero fill (A/B), digits counted
# while the expansion waits for its look-back -- full GPU suite, bench, per-launch times
mkdir -p gpurun_out/r2c23 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c23
timeout 900 python -m pytest tests -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run base
run nooverlap B2S_BWD_OVERLAP=0
run base2
B="python bench.py --gpus 1 --steps 30 --warmup 10 --no-extras --no-e2e --no-cpu-baseline --workload config3_init"
run init_base
run init_nooverlap B2S_BWD_OVERLAP=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("base","nooverlap","base2","init_base","init_nooverlap"):
    try:
        d=json.loads(open(f"gpurun_out/r2c23/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
