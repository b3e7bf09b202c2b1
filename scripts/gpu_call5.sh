#!/bin/bash
# round 2, call 5 (8 GPUs): W=4 / W=8 stripe parity tests, camera-owner + sparse all-reduce tests, scaling benches at N=8 and N=4
mkdir -p gpurun_out/r2c5 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c5
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_stripes.py tests/test_gpu_distributed.py -m gpu -q -rs > $O/pytest_multi_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_multi_gpu.log
run() { # name nproc extra-env...
  name=$1; n=$2; shift 2
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $n --steps 20 --warmup 5 > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; tail -c 400 $O/$name.err | grep -v OMP | tail -3
}
run bench_n8 8 B2S_ROUTE=direct
run bench_n4 4 B2S_ROUTE=direct
run bench_n8_p2p 8 B2S_ROUTE=p2p
python - <<'PY'
import json
for f in ("bench_n8","bench_n4","bench_n8_p2p"):
    try:
        d=json.loads(open(f"gpurun_out/r2c5/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["value"],1), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), {k:v["ms"] for k,v in d["stages"].items()})
        for k,v in d["extras"].items(): print("  ", k, v.get("ms_per_step"), v.get("stage_ms"), v.get("error"))
    except Exception as e:
        print(f, "ERR", e)
PY
