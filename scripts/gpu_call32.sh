#!/bin/bash
# round 2, call 32 (1 GPU): full GPU suite on the final commit of the round
mkdir -p gpurun_out/r2c32 && cd $GRAFT_REPO_ROOT
timeout 100 python -m pytest tests -m gpu -q -rs -x > gpurun_out/r2c32/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c32/pytest_gpu.log
