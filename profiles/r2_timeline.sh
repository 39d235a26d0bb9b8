#!/bin/bash
# sparse-backward tests + kernel timelines of one graph replay (both workloads, list path on / off)
TAG=${1:-tl}
python -m pytest tests/test_gpu_sparse_bwd.py -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log; grep -n "^E  " gpurun_out/${TAG}_tests.log | head -8
for w in smooth scene; do
  python profiles/timeline.py $w gpurun_out/${TAG}_${w}.json > gpurun_out/${TAG}_${w}.txt 2> gpurun_out/${TAG}_${w}.err || tail -3 gpurun_out/${TAG}_${w}.err
  python profiles/timeline.py $w gpurun_out/${TAG}_${w}_nolist.json photo_list_max=-1 > gpurun_out/${TAG}_${w}_nolist.txt 2> gpurun_out/${TAG}_${w}_nolist.err || tail -3 gpurun_out/${TAG}_${w}_nolist.err
  head -1 gpurun_out/${TAG}_${w}.txt | cut -c1-700; head -1 gpurun_out/${TAG}_${w}_nolist.txt | cut -c1-700
done
