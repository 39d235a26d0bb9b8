#!/bin/bash
# Round-2 evidence run on one B200: parity tests (+ report), the bench line (own arm + reference arm), the ncu launch list and
# one `--set full` capture of every hot kernel ON THE SCENE WORKLOAD.
# Usage (repo root): gpurun --timeout 2400 -- 'bash profiles/evidence_r2.sh TAG'; then here: bash profiles/make_summaries_r2.sh TAG
TAG=${1:-r2}
rm -f gpurun_out/${TAG}_parity.jsonl
TDL_PARITY_REPORT=gpurun_out/${TAG}_parity.jsonl python -m pytest tests -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1; tail -1 gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-train --no-bf16"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "list rc=$?"
# scene-workload launches start after the smooth workload's 13 steps (3 eager + 2 profiled + 3 warm + 5 replays): skip 16 per kernel
ncu --set full --clock-control none --import-source on \
    -k regex:"photo_bwd_kernel|photo_warp_kernel|photo_score2_kernel|feat_fwd_nhwc|feat_bwd_nhwc|feat_gather_nhwc|smooth_fwd_kernel|smooth_bwd_kernel" -s 128 -c 8 \
    -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
# the two input-pipeline kernels, TripleD item shape
bash profiles/input_ncu.sh ${TAG}_input
