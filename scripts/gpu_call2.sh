#!/bin/bash
# round 2, call 2 (1 GPU): TMA path tests + A/B timing, remaining trainer tests
mkdir -p gpurun_out/r2c2 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c2
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "tma or projection or forward_backward or golden" > $O/pytest_tma.log 2>&1; echo "pytest tma rc=$?"; tail -3 $O/pytest_tma.log
timeout 900 python -m pytest tests -m gpu -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
for t in 1 0; do
  B2S_PROJECT_TMA=$t timeout 300 python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline > $O/bench_tma$t.json 2> $O/bench_tma$t.err; echo "bench tma=$t rc=$?"
done
python - <<'PY'
import json
for t in (1,0):
    try:
        d=json.loads(open(f"gpurun_out/r2c2/bench_tma{t}.json").read().strip().splitlines()[-1])
        print("tma",t, round(d["ms_per_step"],4), d["roofline"]["frac"], {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e: print(t,"ERR",e)
PY
