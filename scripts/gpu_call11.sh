#!/bin/bash
# round 2, call 11 (1 GPU): thin zero-fill kernel: standalone bandwidth, and the step with / without the side-stream prefill
mkdir -p gpurun_out/r2c11 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c11
python - > $O/fill_bw.txt 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, ".")
from pipeline_pointcloud_b200 import _lib as L
t = torch.empty(278_400_000, device="cuda")      # 1.11 GB, the v_sh of config 3
lib = L.lib(); st = L.stream_ptr(t.device)
for _ in range(3): lib.b2s_zero_fill(L.ptr(t), t.numel() * 4, st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): lib.b2s_zero_fill(L.ptr(t), t.numel() * 4, st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"b2s_zero_fill 1.1136 GB: {ms:.4f} ms = {t.numel()*4/ms/1e6:.0f} GB/s")
PY
cat $O/fill_bw.txt
B="python bench.py --gpus 1 --steps 50 --warmup 10 --no-extras --no-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 300 $B > $O/$name.json 2> $O/$name.err; echo "$name rc=$?"; }
run base
run noprefill B2S_PREFILL=0
run base2
run noprefill2 B2S_PREFILL=0
python - <<'PY'
import json
for f in ("base","noprefill","base2","noprefill2"):
    try:
        d=json.loads(open(f"gpurun_out/r2c11/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],4), round(d["ms_median"],4), {k:v["ms"] for k,v in d["stages"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
