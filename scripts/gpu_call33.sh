#!/bin/bash
# round 2, call 33 (1 GPU, the round's last GPU seconds): the native PLY -> SPZ entry point and the converter command line
mkdir -p gpurun_out/r2c33 && cd $GRAFT_REPO_ROOT
timeout 38 python -m pytest tests/test_gpu_spz.py -m gpu -q -x > gpurun_out/r2c33/pytest_spz.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2c33/pytest_spz.log
