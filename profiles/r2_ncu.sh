#!/bin/bash
# one `ncu --set full` capture of selected kernels of the bench's scene workload: bash profiles/r2_ncu.sh TAG "regex" [skip] [count]
TAG=$1; RX=$2; SKIP=${3:-0}; CNT=${4:-4}
CMD="python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-train"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu.log
