#!/bin/bash
# round 2, call 16 (1 GPU): fused binning v2 (rectangles as a second sort payload from the first pass, direct emission) -- parity tests, A/B against the
# unfused kernels, ranking variants of the onesweep pass, host-enqueue diagnostic, per-launch times
mkdir -p gpurun_out/r2c16 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c16
timeout 600 python -m pytest tests/test_gpu_binning_fused.py tests/test_gpu_spz.py tests/test_gpu_stripes.py -m gpu -q -rs -x > $O/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_new.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run fused
run unfused B2S_BINNING_FUSED=0
run rank0 B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_rank0.so
run rank0_unfused B2S_LIB=$PWD/pipeline_pointcloud_b200/libb2s_rank0.so B2S_BINNING_FUSED=0
run fused2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for f in ("fused","unfused","rank0","rank0_unfused","fused2"):
    try:
        d=json.loads(open(f"gpurun_out/r2c16/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
