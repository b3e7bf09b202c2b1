#!/bin/bash
# round 2, call 17 (1 GPU): fused binning v3 (packed rectangles stored by row from the first sort pass, one-word gather in the one-kernel expansion) -- parity tests, A/B against the
# unfused kernels, ranking variants of the onesweep pass, host-enqueue diagnostic, per-launch times
mkdir -p gpurun_out/r2c17 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c17
timeout 600 python -m pytest tests/test_gpu_binning_fused.py tests/test_gpu_spz.py tests/test_gpu_stripes.py -m gpu -q -rs -x > $O/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_new.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run fused
run unfused B2S_BINNING_FUSED=0
run fused2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > $O/ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
    try:
        d=json.loads(open(f"gpurun_out/r2c17/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
