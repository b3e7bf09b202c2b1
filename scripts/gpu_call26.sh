#!/bin/bash
# round 2, call 26 (1 GPU): diagnostic of the packed=True + meta-read step (8.5 ms outliers in call 25)
mkdir -p gpurun_out/r2c26 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c26
timeout 300 python scripts/packed_meta_probe.py > $O/probe.txt 2>&1; echo "probe rc=$?"
B2S_ASYNC_COUNTS=0 timeout 300 python scripts/packed_meta_probe.py > $O/probe_sync.txt 2>&1; echo "probe_sync rc=$?"
cat $O/probe.txt; echo ----; cat $O/probe_sync.txt
