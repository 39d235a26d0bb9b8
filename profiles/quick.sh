#!/bin/bash
# quick GPU check used during development: feature/photo parity tests + a short bench, per-kernel times
python -m pytest tests -m gpu -x -q > gpurun_out/t.log 2>&1; tail -2 gpurun_out/t.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train > gpurun_out/bq.json 2> gpurun_out/bq.err
python -c "
import json; d=json.load(open('gpurun_out/bq.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value']); [print(' ', k, v['us_per_step']) for k,v in d['roofline']['kernels'].items()]"
