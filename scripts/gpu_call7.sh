#!/bin/bash
# round 2, call 7 (1 GPU): ncu launch list (time + DRAM bytes) and --set full of the hot kernels, after the same command exited 0 without ncu
mkdir -p gpurun_out/r2c7 && cd $GRAFT_REPO_ROOT
O=gpurun_out/r2c7
CMD="python bench.py --gpus 1 --steps 2 --warmup 3 --no-extras --no-e2e --no-cpu-baseline"
$CMD > $O/plain.json 2> $O/plain.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 120 --csv --log-file $O/launches.csv $CMD > $O/ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'composite|project|sh_|coarse_rects|coarse_emit|onesweep|classify' -s 61 -c 16 -o $O/full $CMD > $O/ncu_full.log 2>&1
echo "full rc=$?"; ls -la $O; tail -3 $O/ncu_full.log
