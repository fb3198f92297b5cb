#!/bin/bash
# short device-resident runs of every kernel family with the default tile choice (extra bench.py flags may be appended),
# then the filter-kernel probes (PFDTD_DEBUG_DIF_KEEP: which population of boundary voxels costs what)
#   usage: bash tools/dif_probe.sh [tests] [quick] [bench flags...]
run() { timeout 300 python bench.py "$@" --steps 100 --warmup 10 --no-variants --no-e2e --no-cpu-baseline --no-like-for-like --c4 off | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['config']['kernel'], round(d['roofline']['achieved']), round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms_per_launch']*1e3,1), 'us')"; }
if [ "$1" = "tests" ]; then shift; timeout 900 python -m pytest tests/test_gpu_dif.py tests/test_gpu_interp.py tests/test_gpu_large_grids.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3; fi
ORD="0 2"; if [ "$1" = "quick" ]; then shift; ORD="2"; fi
for d in f32 f64; do for u in 0 2 3; do for o in $ORD; do
  echo "== $d update-type $u dif $o"; run --dtype $d --update-type $u --dif-order $o "$@"
done; done; done
for d in f32; do for k in 0 1 2; do echo "== $d dif2 keep=$k"; PFDTD_DEBUG_DIF_KEEP=$k run --dif-order 2 --dtype $d "$@"; done; done
