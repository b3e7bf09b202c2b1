#!/bin/bash
# round 2, call 24 (2 GPUs): multi-GPU tests with the fused binning, the new compositing loops and the side-stream counts; N=2 bench
mkdir -p gpurun_out/r2c24 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c24
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_stripes.py tests/test_gpu_trainer.py -m gpu -q -rs -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 30 --warmup 5 --no-e2e > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c24/bench_n2.json").read().strip().splitlines()[-1])
print("n2", round(d["ms_per_step"],3), round(d["value"],1), {k:v["ms"] for k,v in d["stages"].items()})
for k,v in d["extras"].items(): print("  ", k, v.get("ms_per_step"), v.get("stage_ms"), v.get("error"))
PY
