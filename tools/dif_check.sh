#!/bin/bash
# GPU: filter-boundary parity tests, then short device-resident benches of the filter variants
timeout 600 python -m pytest tests/test_gpu_dif.py tests/test_gpu_interp.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for a in "--dif-order 2" "--dif-order 2 --dtype f64" "--dif-order 2 --update-type 3" "--dif-order 1" "--dif-order 4" "--dif-order 2 --update-type 2" "$@"; do
  timeout 300 python bench.py $a --steps 300 --no-variants --no-e2e --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$a', round(d['value']), d['config']['kernel'], round(d['roofline']['achieved']), d['roofline']['kernel_ms_per_step'])"
done
