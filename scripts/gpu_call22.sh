#!/bin/bash
# round 2, call 22 (1 GPU): binning tests with the final look-back structure, bench, ncu --set full of the binning and compositing
# kernels of one steady-state step (after the same command exited 0 without ncu), launch list
mkdir -p gpurun_out/r2c22 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c22
timeout 900 python -m pytest tests/test_gpu_binning_fused.py tests/test_gpu_parity.py -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
timeout 300 $B > $O/base.json 2> $O/base.err; echo "base rc=$?"
P="python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
timeout 300 $P > $O/plain.json 2> $O/plain.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'coarse_emit|onesweep|hist_kernel|composite|classify' -s 40 -c 11 -o $O/full -f $P > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $P > $O/ncu_l.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c22/base.json").read().strip().splitlines()[-1])
print("base", round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
PY
ls -la $O
