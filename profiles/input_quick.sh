#!/bin/bash
# input-pipeline development check: bit-exactness tests + device time of tdl_input_fwd on the TripleD item shape
python -m pytest tests/test_input_pipeline.py -q -x 2>&1 | tail -3
python - <<'PY'
import importlib, sys, torch
sys.path.insert(0, '.')
import bench
with torch.cuda.stream(torch.cuda.Stream()):
    print(bench.input_pipeline_bench(8, 192, 640, torch.device("cuda:0"), 50, cpu_baseline=False))
PY
