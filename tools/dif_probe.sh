#!/bin/bash
# short device-resident runs of every kernel family with the default tile choice (extra bench.py flags may be appended)
run() { timeout 300 python bench.py "$@" --steps 300 --no-variants --no-e2e --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['config']['kernel'], round(d['roofline']['achieved']), d['roofline']['kernel_ms_per_step'])"; }
for d in f32 f64; do for u in 0 2 3; do for o in 0 2; do
  echo "== $d update-type $u dif $o"; run --dtype $d --update-type $u --dif-order $o "$@"
done; done; done
