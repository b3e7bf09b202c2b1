#!/bin/bash
# round 2, call 12 (1 GPU): composite_bwd direct-atomic path for (warp, splat) pairs with few contributing pixels: A/B of the threshold
mkdir -p gpurun_out/r2c12 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c12
P=$PWD/pipeline_pointcloud_b200
B2S_LIB=$P/libb2s_d4.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -m gpu -q -x > $O/pytest_d4.log 2>&1; echo "pytest d4 rc=$?"; tail -3 $O/pytest_d4.log
run() { name=$1; wl=$2; shift 2; env "$@" timeout 300 python bench.py --gpus 1 --steps 40 --warmup 10 --no-extras --no-e2e --no-cpu-baseline --workload $wl > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
for v in d0 d2 d4 d8; do
  if [ $v = d0 ]; then L=$P/libb2s.so; else L=$P/libb2s_$v.so; fi
  run $v config3 B2S_LIB=$L
  run init_$v config3_init B2S_LIB=$L
done
python - <<'PY'
import json
for f in ("d0","d2","d4","d8","init_d0","init_d2","init_d4","init_d8"):
    try:
        d=json.loads(open(f"gpurun_out/r2c12/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items() if k.startswith("composite")})
    except Exception as e:
        print(f, "ERR", e)
PY
