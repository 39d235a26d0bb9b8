#!/bin/bash
# make_summaries_r2.sh TAG: turns gpurun_out/TAG_{launches.csv,full.ncu-rep,parity.jsonl} (profiles/evidence_r2.sh) into the committed
# profiles/TAG_launches.csv, TAG_launches_summary.md, TAG_ncu_full_summary.md, profiles/parity_r2.md and profiles/ncu_traffic.json.
set -e
TAG=$1; cd "$(dirname "$0")/.."
PKG=$(ls -d tripled*_b200); TMP=$(mktemp -d); (cd $TMP && cuobjdump -xelf all $OLDPWD/$PKG/libtdl.so > /dev/null)
grep -v "^==" gpurun_out/${TAG}_launches.csv | cut -d, -f1-2,5,9- > /dev/null 2>&1 || true
python - "$TAG" <<'PY'
import csv, sys
tag = sys.argv[1]
rows = [l for l in open(f'gpurun_out/{tag}_launches.csv') if not l.startswith('==')]
# keep our kernels (namespace tdl) plus a one-line count of everything else, to keep the committed list small
keep = [rows[0]] + [l for l in rows[1:] if 'tdl::' in l]
open(f'profiles/{tag}_launches.csv', 'w').writelines(keep)
print(len(rows) - 1, 'launches in the capture,', len(keep) - 1, 'of them tdl:: kernels')
PY
{ echo "# Round 2 ($TAG) -- ncu launch list"; echo
  echo 'Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --steps 2 --warmup 1 --repeats 1 --no-cpu-baseline --no-train --no-bf16` (`profiles/evidence_r2.sh`)'
  echo "(both workloads: eager warm-up + event-profiling pass + graph replays of the loss step, then the e2e loops; cold-cache, serialised: compare SHARES)."
  echo "Raw list of the library's kernels: \`profiles/${TAG}_launches.csv\`; produced with \`profiles/launch_summary.py\`."; echo
  python profiles/launch_summary.py gpurun_out/${TAG}_launches.csv | grep -E "kernel|---|tdl::" | head -30; } > profiles/${TAG}_launches_summary.md
REP=gpurun_out/${TAG}_full.ncu-rep
{ echo "# Round 2 ($TAG) -- \`ncu --set full --clock-control none --import-source on\`, scene workload"; echo
  echo 'Command: see `profiles/evidence_r2.sh` (one launch of every hot kernel of the bench default workload B=8, 192x640, S=2, 4 scales, C=64 NHWC features, "scene" frames).'
  echo "Per-launch values, produced with \`profiles/ncu_summary.py\`."; echo
  python profiles/ncu_summary.py $REP
  echo; echo "## Speed-of-light section per kernel (\`ncu --page details\`)"; echo; echo '```'
  ncu -i $REP --page details 2>/dev/null | grep -E "^  [a-z_A-Z].*\(|L1/TEX Cache Throughput|Issue Slots Busy|DRAM Throughput|Duration  |L2 Cache Throughput|Mem Pipes Busy|Achieved Occupancy|Theoretical Occupancy|Registers Per Thread"
  echo '```'
  for k in "photo_score2:photo_score2_kernelILi2E:tdl_photo2" "photo_bwd_kernel:photo_bwd_kernelILi2ELb1E:tdl_photo" "feat_fwd_nhwc:feat_fwd_nhwc_bulk_kernelILi2EfE:tdl_feat2"; do
    IFS=: read rx mang cub <<< "$k"
    echo; echo "## \`$rx\`: SASS opcode histogram (\`profiles/sass_opcount.py\`)"; echo; echo '```'
    python profiles/sass_opcount.py $REP $rx 2>/dev/null | head -14; echo '```'
    echo; echo "## \`$rx\`: executed instructions per source line (\`profiles/line_profile.py\`, top 12)"; echo; echo '```'
    python profiles/line_profile.py $REP $rx $TMP/$cub.sm_100a.cubin $mang 12 2>/dev/null | cut -c1-170; echo '```'
  done
  if [ -f gpurun_out/${TAG}_input.ncu-rep ]; then
    echo; echo "## Input pipeline kernels (\`profiles/input_ncu.sh\`: 8 items x 3 frames at 192x640, jitter on every image, 16 erase boxes)"; echo
    python profiles/ncu_summary.py gpurun_out/${TAG}_input.ncu-rep
  fi; } > profiles/${TAG}_ncu_full_summary.md
python - "$REP" "$TAG" <<'PY'
import csv, json, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
names = {'photo_warp_kernel': 'photo_warp', 'photo_score2_kernel': 'photo_score', 'feat_fwd_nhwc_kernel': 'feat_fwd', 'feat_fwd_nhwc_bulk_kernel': 'feat_fwd',
         'feat_bwd_nhwc_kernel': 'feat_bwd', 'photo_bwd_kernel': 'photo_bwd', 'feat_gather_nhwc_kernel': 'feat_gather',
         'smooth_fwd_kernel': 'smooth_fwd', 'smooth_bwd_kernel': 'smooth_bwd'}
def tob(col, r):
    v = float(r[idx[col]].replace(',', ''))
    return int(v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[rows[1][idx[col]]])
k = {}
for r in rows[2:]:
    n = r[idx['Kernel Name']].split('(')[0].replace('void ', '').split('<')[0]
    if n in names:
        k[names[n]] = tob('dram__bytes_read.sum', r) + tob('dram__bytes_write.sum', r)
json.dump({"round": 2, "source": f"profiles/{tag}_ncu_full_summary.md (ncu --set full, one launch each, scene workload)",
           "workload": [8, 192, 640, 2, 64], "kernels": k, "scene": {"kernels": k}}, open('profiles/ncu_traffic.json', 'w'), indent=1)
print(k)
PY
python tests/parity_report_md.py gpurun_out/${TAG}_parity.jsonl > profiles/parity_r2.md
rm -rf $TMP
