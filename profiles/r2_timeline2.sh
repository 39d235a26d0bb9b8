#!/bin/bash
# parity tests of the photometric path + kernel timelines of one graph replay, default library and variants/*.so
TAG=${1:-tl}; WL=${2:-smooth}
python -m pytest tests/test_gpu_sparse_bwd.py tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_bench_configs.py -q -x > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log; grep -n "^E  " gpurun_out/${TAG}_tests.log | head -8
for w in smooth scene; do
  python profiles/timeline.py $w gpurun_out/${TAG}_${w}.json > gpurun_out/${TAG}_${w}.txt 2> gpurun_out/${TAG}_${w}.err || tail -3 gpurun_out/${TAG}_${w}.err
  head -1 gpurun_out/${TAG}_${w}.txt | cut -c1-400; grep "feat\|photo\|smooth" gpurun_out/${TAG}_${w}.txt
done
for v in variants/*.so; do
  [ -f "$v" ] || continue
  n=$(basename "$v" .so)
  TDL_LIB_PATH="$PWD/$v" python profiles/timeline.py $WL > "gpurun_out/${TAG}_${n}.txt" 2> "gpurun_out/${TAG}_${n}.err" || tail -3 "gpurun_out/${TAG}_${n}.err"
  echo "$n"; head -1 "gpurun_out/${TAG}_${n}.txt" | cut -c1-200; grep "feat\|photo\|smooth" "gpurun_out/${TAG}_${n}.txt"
done
