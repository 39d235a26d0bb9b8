#!/bin/bash
# variants.sh <name>...: short bench per build/variants/libtdl_<name>.so, prints the per-kernel times that match $KFILTER
for v in "$@"; do
  TDL_LIB_PATH=/root/repo/build/variants/libtdl_$v.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err
  python -c "
import json,os,re; d=json.load(open('gpurun_out/v_$v.json')); f=os.environ.get('KFILTER','.')
print('$v', d['ms_per_step'], ' '.join('%s=%.1f'%(k,v['us_per_step']) for k,v in d['roofline']['kernels'].items() if re.search(f,k)))"
done
