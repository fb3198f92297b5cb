#!/bin/bash
# short device-resident runs of every kernel family with the default tile choice (extra bench.py flags may be appended),
# then the filter-kernel probes (PFDTD_DEBUG_DIF_KEEP: which population of boundary voxels costs what)
run() { timeout 300 python bench.py "$@" --steps 300 --no-variants --no-e2e --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['config']['kernel'], round(d['roofline']['achieved']), d['roofline']['kernel_ms_per_step'])"; }
if [ "$1" = "tests" ]; then shift; timeout 900 python -m pytest tests/test_gpu_dif.py tests/test_gpu_interp.py tests/test_gpu_large_grids.py -x -q -m gpu 2>&1 | tail -3; fi
for d in f32 f64; do for u in 0 2 3; do for o in 0 2; do
  echo "== $d update-type $u dif $o"; run --dtype $d --update-type $u --dif-order $o "$@"
done; done; done
for d in f32 f64; do for k in 0 1 2; do echo "== $d dif2 keep=$k"; PFDTD_DEBUG_DIF_KEEP=$k run --dif-order 2 --dtype $d "$@"; done; done
