#!/bin/bash
# Round-2 development run on one B200: usage (repo root): gpurun --timeout 1500 -- 'bash profiles/r2_run.sh TAG [tests|notests] [bench args...]'
TAG=${1:-r2}; MODE=${2:-tests}; shift 2
if [ "$MODE" = tests ]; then
  rm -f gpurun_out/${TAG}_parity.jsonl
  TDL_PARITY_REPORT=gpurun_out/${TAG}_parity.jsonl python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_tests.log 2>&1; tail -15 gpurun_out/${TAG}_tests.log
fi
python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/${TAG}_bench.json'))
    print('value', d['value'], 'ms', d['ms_per_step'], d['repeats'], 'e2e', d['e2e']['value'], 'io', d['e2e_images_only']['value'])
    for w in ('smooth', 'scene'):
        b = d.get(w)
        if b:
            print(w, b['images_per_s'], b['ms_per_step'], 'ident', b['identity_frac'])
            for k, v in b['kernels'].items():
                print('   ', k, v['us_per_step'], v.get('gbs'))
    print('roofline', {k: v for k, v in d['roofline'].items() if k != 'kernels'})
    print('train', d.get('train_step'))
    print('cpu', d.get('cpu_baseline'))
except Exception as e:
    print('no bench json', e)
PY
