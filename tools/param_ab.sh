#!/bin/bash
run() { timeout 300 python bench.py "$@" --steps 100 --warmup 10 --no-variants --no-e2e --no-cpu-baseline --no-like-for-like --c4 off | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"; }
for lib in default nct nfp both; do
  if [ "$lib" = "default" ]; then unset PFDTD_LIB_PATH; else export PFDTD_LIB_PATH=$PWD/build/variants/libpfdtd_b200_$lib.so; fi
  echo -n "lib=$lib f64 iiso dif2: "; run --dtype f64 --update-type 3 --dif-order 2
  echo -n "lib=$lib f32 iiso dif2: "; run --dtype f32 --update-type 3 --dif-order 2
  echo -n "lib=$lib f64 fwd dif2: "; run --dtype f64 --update-type 0 --dif-order 2
done
