#!/bin/bash
# round 2, call 20 (1 GPU): privatised global histograms (16 copies), histogram kernel with 4 consecutive rows per thread (A/B
# against 1), persistent pinned ring for the counts -- GPU suite, bench, forward-only timing, per-launch times
mkdir -p gpurun_out/r2c20 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c20
timeout 900 python -m pytest tests -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run base
run u1 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_u1.so
run base2
timeout 300 python scripts/host_overhead.py > $O/host_overhead.txt 2>&1; echo "host_overhead rc=$?"; head -3 $O/host_overhead.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("base","u1","base2"):
    try:
        d=json.loads(open(f"gpurun_out/r2c20/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
