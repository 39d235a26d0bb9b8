#!/bin/bash
# Round-end evidence run on one B200: parity tests, the bench line (own arm + reference arm), the ncu launch list and one
# `--set full` capture per hot kernel.  Usage (from the repo root): gpurun --timeout 1500 -- 'bash profiles/evidence.sh TAG'
TAG=${1:-r1}
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; tail -1 gpurun_out/${TAG}_tests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on \
    -k regex:"photo_bwd|photo_warp|photo_score|feat_fwd|feat_bwd|feat_gather|feat_overflow" -s 14 -c 7 \
    -o gpurun_out/${TAG}_full python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
