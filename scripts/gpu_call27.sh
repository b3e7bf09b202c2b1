#!/bin/bash
# round 2, call 27 (1 GPU): the same probe after the weak-reference fix of the packed-meta hook
mkdir -p gpurun_out/r2c27 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c27
timeout 300 python scripts/packed_meta_probe.py > $O/probe.txt 2>&1; echo "probe rc=$?"
cat $O/probe.txt
