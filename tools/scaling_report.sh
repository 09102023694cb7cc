#!/bin/bash
# GPU-count scaling report in the reference's scaling-output.txt format (struct3d_b200/lib/scaling):
#   tools/scaling_report.sh <control file> <perc file> <ref GPUs> <GPU counts...>
# runs the reference count first (writes the header), then every listed count (appends its lines).
CTRL=$1; PERC=$2; REF=$3; shift 3
run() {
  if [ "$1" = 1 ]; then struct3d_b200/lib/scaling "$CTRL" -options_file "$PERC"
  else python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port 29541 struct3d_b200/lib/scaling "$CTRL" -options_file "$PERC"; fi
}
run "$REF" || exit 1
for n in "$@"; do [ "$n" = "$REF" ] && continue; run "$n" || exit 1; done
OUT=$(grep openmptest_output_file "$PERC" | awk '{print $2}')
cat "$OUT"
