#!/bin/bash
# round 2, call 8 (1 GPU): compute-sanitizer memcheck on a small scene that runs every single-GPU kernel of the path
mkdir -p gpurun_out/r2c8 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c8
timeout 300 python scripts/sanitize_scene.py > $O/plain.log 2>&1; echo "plain rc=$?"; tail -3 $O/plain.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_scene.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -8 $O/memcheck.log
