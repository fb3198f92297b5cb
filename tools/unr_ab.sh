#!/bin/bash
# A/B of plane-loop unroll factors (tools/build_variant.py -DPFDTD_UNR_*): bash tools/unr_ab.sh
run() { timeout 300 python bench.py "$@" --steps 100 --warmup 10 --no-variants --no-e2e --no-cpu-baseline --no-like-for-like --c4 off | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), round(d['roofline']['kernel_ms_per_launch']*1e3,1), 'us')"; }
for lib in default f64u1 allu1; do
  if [ "$lib" = "default" ]; then unset PFDTD_LIB_PATH; else export PFDTD_LIB_PATH=$PWD/build/variants/libpfdtd_b200_$lib.so; fi
  echo -n "lib=$lib f64 forward dif2: "; run --dtype f64 --update-type 0 --dif-order 2
  echo -n "lib=$lib f64 centred dif2: "; run --dtype f64 --update-type 2 --dif-order 2
  if [ "$lib" != "f64u1" ]; then
  echo -n "lib=$lib f32 forward dif0: "; run --update-type 0 --dif-order 0
  echo -n "lib=$lib f32 centred dif0: "; run --update-type 2 --dif-order 0
  echo -n "lib=$lib f64 forward dif0: "; run --dtype f64 --update-type 0 --dif-order 0
  fi
done
unset PFDTD_LIB_PATH
echo -n "default f32 forward dif2: "; run --update-type 0 --dif-order 2
echo -n "default f32 centred dif2: "; run --update-type 2 --dif-order 2
echo -n "default f64 iiso dif2: "; run --dtype f64 --update-type 3 --dif-order 2
echo -n "default f32 iiso dif2: "; run --dtype f32 --update-type 3 --dif-order 2
python -m pytest tests/test_gpu_interp.py tests/test_gpu_dif.py tests/test_gpu_large_grids.py tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -2
