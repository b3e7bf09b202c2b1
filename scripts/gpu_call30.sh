#!/bin/bash
# round 2, call 30 (1 GPU): the round's record run -- full GPU suite, smoke(), the driver's bench command (default flags) and the
# reference arm, ncu --set full of one steady-state step (after the same command exited 0 without ncu), launch list with DRAM bytes
mkdir -p gpurun_out/r2c30 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c30
timeout 900 python -m pytest tests -m gpu -q -rs > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
P="python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
timeout 300 $P > $O/plain.json 2> $O/plain.err; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'project_fwd|hist_kernel|onesweep|coarse_|composite|classify|sh_|project_bwd|zero_fill' -s 84 -c 21 -o $O/full -f $P > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches.csv $P > $O/ncu_l.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import json
for f in ("bench_default","bench_reference"):
    try:
        d=json.loads(open(f"gpurun_out/r2c30/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("ms_per_step"), d.get("ms_median"), d.get("value"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
        if "stages" in d: print("   ", {k:v["ms"] for k,v in d["stages"].items()})
        for k,v in (d.get("extras") or {}).items(): print("   ", k, v.get("ms_per_step"), v.get("ms_median"), v.get("error"))
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la $O | head -20
