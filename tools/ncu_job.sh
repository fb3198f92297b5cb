#!/bin/bash
# ncu --set full captures of the update kernel for the listed bench variants, exported to CSV on the box (the .ncu-rep files
# are too large to travel back): raw metrics + per-source-line page.   usage: bash tools/ncu_job.sh <outdir> "<bench args>" ...
OUT=$1; shift
B="--no-cpu-baseline --no-variants --no-like-for-like --c4 off --no-e2e"
mkdir -p $OUT
for a in "$@"; do
  n=$(echo "prof $a" | tr -d ' ' | tr '-' '_')
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fdtd_update -s 4 -c 1 -f -o /tmp/$n python bench.py $a --steps 6 --warmup 3 $B > $OUT/$n.log 2>&1; echo "ncu [$a] rc=$?"
  ncu -i /tmp/$n.ncu-rep --page raw --csv > $OUT/$n.raw.csv 2>/dev/null
  ncu -i /tmp/$n.ncu-rep --page source --csv > $OUT/$n.source.csv 2>/dev/null
  ls -la /tmp/$n.ncu-rep $OUT/$n.*.csv
done
