#!/bin/bash
# round 2, call 21 (1 GPU): look-back restructured -- onesweep: aggregate published early, local reorder before the look-back,
# 32-wide window (A/B 8 / 16); coarse-bin expansion: CTA-wide (256-wide) look-back -- binning tests, bench, per-launch times
mkdir -p gpurun_out/r2c21 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c21
timeout 900 python -m pytest tests/test_gpu_binning_fused.py tests/test_gpu_parity.py tests/test_gpu_stripes.py -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run base
run lb8 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_lb8.so
run lb16 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_lb16.so
run base2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("base","lb8","lb16","base2"):
    try:
        d=json.loads(open(f"gpurun_out/r2c21/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
