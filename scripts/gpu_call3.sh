#!/bin/bash
# round 2, call 3 (2 GPUs): direct route (projection fused with the exchange) -- tests, then N=2 benches per route
mkdir -p gpurun_out/r2c3 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c3
timeout 1200 python -m pytest tests -m gpu -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
for route in direct p2p; do
  B2S_ROUTE=$route timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e > $O/bench_n2_$route.json 2> $O/bench_n2_$route.err; echo "bench2 $route rc=$?"
  tail -c 600 $O/bench_n2_$route.err
done
python - <<'PY'
import json
for f in ("bench_n2_direct","bench_n2_p2p"):
    try:
        d=json.loads(open(f"gpurun_out/r2c3/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["value"],1), {k:v["ms"] for k,v in d["stages"].items()})
        for k,v in d["extras"].items(): print("  ", k, v.get("ms_per_step"), v.get("stage_ms"), v.get("error"))
    except Exception as e:
        print(f, "ERR", e)
PY
