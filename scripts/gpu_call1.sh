#!/bin/bash
# round 2, call 1 (2 GPUs): every GPU test incl. the >=2-GPU ones, bench at N=1 and N=2
mkdir -p gpurun_out/r2c1 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c1
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench1 rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench2 rc=$?"
tail -c 1500 $O/bench_n1.err; tail -c 1500 $O/bench_n2.err
python - <<'PY'
import json
for f in ("bench_n1","bench_n2"):
    try:
        d=json.loads(open(f"gpurun_out/r2c1/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), round(d["value"],1), {k:v["ms"] for k,v in d["stages"].items()})
        print(" extras", {k:(round(v.get("ms_per_step",0),3) if "ms_per_step" in v else v) for k,v in d["extras"].items()})
        print(" parity", d.get("parity_check"))
    except Exception as e:
        print(f, "ERR", e)
PY
