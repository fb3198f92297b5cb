#!/bin/bash
# A/B of filter-kernel builds (tools/build_variant.py) and TMA hint bits: bash tools/dif_ab.sh "<lib names>" "<hint values>"
run() { timeout 300 python bench.py "$@" --steps 100 --warmup 10 --no-variants --no-e2e --no-cpu-baseline --no-like-for-like --c4 off | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['config']['kernel'], round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms_per_launch']*1e3,1), 'us')"; }
LIBS=$1; HINTS=$2
for lib in $LIBS; do for h in $HINTS; do
  if [ "$lib" = "default" ]; then unset PFDTD_LIB_PATH; else export PFDTD_LIB_PATH=$PWD/build/variants/libpfdtd_b200_$lib.so; fi
  for cfg in "f32 0" "f32 2" "f32 3" "f64 0" "f64 3"; do dt=${cfg% *}; ut=${cfg#* }
    echo -n "lib=$lib hints=$h $dt ut=$ut dif2: "; run --dtype $dt --update-type $ut --dif-order 2 --tma-hints $h
  done
  echo -n "lib=$lib hints=$h f32 ut=0 dif2 keep=0: "; PFDTD_DEBUG_DIF_KEEP=0 run --dif-order 2 --tma-hints $h
  echo -n "lib=$lib hints=$h f32 ut=0 dif0: "; run --dif-order 0 --tma-hints $h
  echo -n "lib=$lib hints=$h f64 ut=0 dif0: "; run --dtype f64 --dif-order 0 --tma-hints $h
done; done
