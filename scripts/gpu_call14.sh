#!/bin/bash
# round 2, call 14 (8 GPUs): W=4 / W=8 stripe parity with the final kernels, scaling benches N=8, 4, 1 on one box
mkdir -p gpurun_out/r2c14 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c14
timeout 600 python -m pytest tests/test_gpu_stripes.py -m gpu -q -rs -k "w4 or w8" > $O/pytest_w4_w8.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_w4_w8.log
run() { name=$1; n=$2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run bench_n8 8
run bench_n4 4
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench_n1 rc=$?"
python - <<'PY'
import json
for f in ("bench_n8","bench_n4","bench_n1"):
    try:
        d=json.loads(open(f"gpurun_out/r2c14/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["value"],1), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), {k:v["ms"] for k,v in d["stages"].items()})
        for k,v in d["extras"].items(): print("  ", k, v.get("ms_per_step"), v.get("stage_ms"), v.get("error"))
    except Exception as e:
        print(f, "ERR", e)
PY
