#!/bin/bash
# feature-kernel tests + parity + kernel timelines (both workloads)
TAG=${1:-tl}
python -m pytest tests/test_gpu_feat_nhwc.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -q -x > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log; grep -n "^E  " gpurun_out/${TAG}_tests.log | head -8
for w in smooth scene; do
  python profiles/timeline.py $w gpurun_out/${TAG}_${w}.json > gpurun_out/${TAG}_${w}.txt 2> gpurun_out/${TAG}_${w}.err || tail -3 gpurun_out/${TAG}_${w}.err
  head -1 gpurun_out/${TAG}_${w}.txt | cut -c1-300; grep "feat\|photo\|smooth" gpurun_out/${TAG}_${w}.txt
done
