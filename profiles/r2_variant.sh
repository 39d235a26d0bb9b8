#!/bin/bash
# bench a library variant: bash profiles/r2_variant.sh TAG /path/to/libtdl_variant.so [bench args]
TAG=$1; LIB=$2; shift 2
TDL_LIB_PATH=$LIB python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_bench.json'))
    for w in ('smooth', 'scene'):
        b = d.get(w)
        if b:
            print('${TAG}', w, b['images_per_s'], b['ms_per_step'], ' | '.join(f"{k} {v['us_per_step']:.0f}" for k, v in b['kernels'].items() if k.startswith('photo')))
except Exception as e:
    print('no bench json', e)
PY
