#!/bin/bash
# round 2, call 4 (2 GPUs): new scan/gather, viewmat gradients, N=2 bench
mkdir -p gpurun_out/r2c4 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c4
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_stripes.py tests/test_gpu_parity.py -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
B2S_ROUTE=direct timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e > $O/bench_n2_direct.json 2> $O/bench_n2_direct.err; echo "bench2 rc=$?"
tail -c 600 $O/bench_n2_direct.err
python - <<'PY'
import json
for f in ("bench_n2_direct",):
    try:
        d=json.loads(open(f"gpurun_out/r2c4/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["value"],1), {k:v["ms"] for k,v in d["stages"].items()})
        for k,v in d["extras"].items(): print("  ", k, v.get("ms_per_step"), v.get("stage_ms"), v.get("error"))
        print(d["roofline"])
    except Exception as e:
        print(f, "ERR", e)
PY
