#!/bin/bash
# round 2, call 6 (2 GPUs): whole GPU suite (lazy packed meta, gather variants, sparse all-reduce), N=1 and N=2 benches
mkdir -p gpurun_out/r2c6 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c6
timeout 1200 python -m pytest tests -m gpu -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench1 rc=$?"; tail -c 300 $O/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 rc=$?"; tail -c 300 $O/bench_n2.err | grep -v OMP
python - <<'PY'
import json
for f in ("bench_n1","bench_n2"):
    try:
        d=json.loads(open(f"gpurun_out/r2c6/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["value"],1), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), {k:v["ms"] for k,v in d["stages"].items()})
        for k,v in d["extras"].items(): print("  ", k, v.get("ms_per_step"), v.get("stage_ms"), v.get("error"))
        print("  parity", (d.get("parity_check") or {}).get("pass"), "roofline", d["roofline"]["frac"])
    except Exception as e:
        print(f, "ERR", e)
PY
