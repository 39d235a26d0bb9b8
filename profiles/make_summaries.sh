#!/bin/bash
# make_summaries.sh TAG: turns gpurun_out/TAG_{launches.csv,full.ncu-rep} (profiles/evidence.sh) into the committed
# profiles/TAG_launches.csv, TAG_launches_summary.md, TAG_ncu_full_summary.md and profiles/ncu_traffic.json.
set -e
TAG=$1; cd "$(dirname "$0")/.."
PKG=$(ls -d tripled*_b200); TMP=$(mktemp -d); (cd $TMP && cuobjdump -xelf all $OLDPWD/$PKG/libtdl.so > /dev/null)
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches.csv
{ echo "# Round 1 ($TAG) -- ncu launch list"; echo
  echo 'Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train` (`profiles/evidence.sh`)'
  echo "(eager warm-up + event-profiling pass + graph warm-up of the loss step; cold-cache, serialised: compare SHARES)."
  echo "Raw list: \`profiles/${TAG}_launches.csv\`; produced with \`profiles/launch_summary.py\`."; echo
  python profiles/launch_summary.py profiles/${TAG}_launches.csv | head -24; } > profiles/${TAG}_launches_summary.md
REP=gpurun_out/${TAG}_full.ncu-rep
{ echo "# Round 1 ($TAG) -- \`ncu --set full --clock-control none --import-source on\`"; echo
  echo 'Command: `ncu --set full --clock-control none --import-source on -k regex:"photo_bwd|photo_warp|photo_score|feat_fwd|feat_bwd|feat_gather|feat_overflow" -s 14 -c 7 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train` (`profiles/evidence.sh`)'
  echo "(bench default workload: B=8, 192x640, S=2, 4 scales, C=64, trainable features).  Per-launch values, produced with \`profiles/ncu_summary.py\`."; echo
  python profiles/ncu_summary.py $REP
  echo; echo "## Speed-of-light section per kernel (\`ncu --page details\`)"; echo; echo '```'
  ncu -i $REP --page details 2>/dev/null | grep -E "^  [a-z_A-Z].*\(|L1/TEX Cache Throughput|Issue Slots Busy|DRAM Throughput|Duration  |L2 Cache Throughput|Mem Pipes Busy|Achieved Occupancy|Theoretical Occupancy|Registers Per Thread"
  echo '```'
  for k in "photo_score:photo_score_kernelILi2E:tdl_photo" "photo_bwd:photo_bwd_kernelILi2ELb1E:tdl_photo" "feat_gather:feat_gather_kernel:tdl_feat" "feat_bwd_bucket:feat_bwd_bucket_kernel:tdl_feat"; do
    IFS=: read rx mang cub <<< "$k"
    echo; echo "## \`$rx\`: SASS opcode histogram (\`profiles/sass_opcount.py\`)"; echo; echo '```'
    python profiles/sass_opcount.py $REP $rx 2>/dev/null | head -14; echo '```'
    echo; echo "## \`$rx\`: executed instructions per source line (\`profiles/line_profile.py\`, top 12)"; echo; echo '```'
    python profiles/line_profile.py $REP $rx $TMP/$cub.sm_100a.cubin $mang 12 2>/dev/null | cut -c1-170; echo '```'
  done; } > profiles/${TAG}_ncu_full_summary.md
python - "$REP" "$TAG" <<'PY'
import csv, json, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
names = {'photo_warp_kernel': 'photo_warp', 'photo_score_kernel': 'photo_score', 'feat_fwd_kernel': 'feat_fwd',
         'feat_bwd_bucket_kernel': 'feat_bwd', 'photo_bwd_kernel': 'photo_bwd', 'feat_gather_kernel': 'feat_gather',
         'feat_overflow_kernel': 'feat_overflow'}
def tob(col, r):
    v = float(r[idx[col]].replace(',', ''))
    return int(v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[rows[1][idx[col]]])
k = {}
for r in rows[2:]:
    n = r[idx['Kernel Name']].split('(')[0].replace('void ', '').split('<')[0]
    k[names[n]] = tob('dram__bytes_read.sum', r) + tob('dram__bytes_write.sum', r)
json.dump({"round": 1, "source": f"profiles/{tag}_ncu_full_summary.md (ncu --set full, one launch each)",
           "workload": [8, 192, 640, 2, 64], "kernels": k}, open('profiles/ncu_traffic.json', 'w'), indent=1)
print(k)
PY
rm -rf $TMP
