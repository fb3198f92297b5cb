#!/bin/bash
run() { timeout 300 python bench.py "$@" --steps 300 --no-variants --no-e2e --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['config']['kernel'], round(d['roofline']['achieved']), d['roofline']['kernel_ms_per_step'])"; }
timeout 600 python -m pytest tests/test_gpu_dif.py tests/test_gpu_interp.py -x -q -m gpu 2>&1 | tail -2
for d in f32 f64; do for k in 0 1 2 3; do echo "== $d dif2 keep=$k"; PFDTD_DEBUG_DIF_KEEP=$k run --dif-order 2 --dtype $d; done; done
echo "== f32 iiso dif2"; run --update-type 3 --dif-order 2
echo "== f32 centred dif2"; run --update-type 2 --dif-order 2
