#!/bin/bash
# ncu --set full of the two input-pipeline kernels (TripleD item shape): bash profiles/input_ncu.sh TAG
TAG=${1:-inp}
cat > /tmp/inp.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
with torch.cuda.stream(torch.cuda.Stream()):
    print(bench.input_pipeline_bench(8, 192, 640, torch.device("cuda:0"), 3, cpu_baseline=False))
PY
python /tmp/inp.py > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"input_" -s 4 -c 2 -o gpurun_out/${TAG} python /tmp/inp.py > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log
