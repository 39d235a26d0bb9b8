#!/bin/bash
# feature-path development check: NHWC parity tests + a short bench
TAG=${1:-fq}; shift
python -m pytest tests/test_gpu_feat_nhwc.py tests/test_gpu_feat_gather.py tests/test_gpu_parity.py tests/test_gpu_bench_configs.py -q -x > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log; grep -n "^E  " gpurun_out/${TAG}_tests.log | head -8
python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${TAG}_bench.err | cut -c1-300
python - <<PY
import json
d = json.load(open('gpurun_out/${TAG}_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'io', d['e2e_images_only']['value'])
for w in ('smooth', 'scene', 'scene_bf16_features'):
    b = d.get(w)
    if b:
        print(w, b['images_per_s'], b['ms_per_step'], ' | '.join(f"{k} {v['us_per_step']:.0f}" for k, v in b['kernels'].items()))
PY
