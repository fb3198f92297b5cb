#!/bin/bash
# One GPU-box session: GPU tests, bench (ours + reference arm), ncu launch list, ncu full captures.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [what...]   what in {tests,smoke,bench,ref,launches,ncu}
TAG=${1:-r01}; shift
WHAT=${@:-tests smoke bench ref launches ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/nvidia_smi.csv 2>&1
nproc > $OUT/nproc.txt
B="--no-cpu-baseline --no-variants"
for w in $WHAT; do
case $w in
tests)
  timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log;;
bench)   # default = BASELINE config 2 with order-2 filter boundaries, fp32; then the other dtype / boundary / scheme with e2e
  timeout 1200 python bench.py > $OUT/bench_f32_dif2.json 2> $OUT/bench_f32_dif2.err; echo "bench rc=$?"; cut -c1-1500 $OUT/bench_f32_dif2.json
  timeout 1200 python bench.py --dtype f64 --steps 500 $B > $OUT/bench_f64_dif2.json 2> $OUT/bench_f64_dif2.err; cut -c1-300 $OUT/bench_f64_dif2.json
  timeout 1200 python bench.py --dif-order 0 $B > $OUT/bench_f32.json 2> $OUT/bench_f32.err; cut -c1-300 $OUT/bench_f32.json
  timeout 1200 python bench.py --dif-order 0 --dtype f64 --steps 500 $B > $OUT/bench_f64.json 2> $OUT/bench_f64.err; cut -c1-300 $OUT/bench_f64.json
  timeout 1200 python bench.py --update-type 3 --steps 500 $B > $OUT/bench_f32_iiso_dif2.json 2> $OUT/bench_f32_iiso_dif2.err; cut -c1-300 $OUT/bench_f32_iiso_dif2.json;;
ref)
  timeout 1200 python bench.py --impl reference > $OUT/bench_ref_f32.json 2> $OUT/bench_ref_f32.err; cut -c1-1200 $OUT/bench_ref_f32.json
  timeout 1200 python bench.py --impl reference --dtype f64 --steps 500 > $OUT/bench_ref_f64.json 2> $OUT/bench_ref_f64.err; cut -c1-400 $OUT/bench_ref_f64.json;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv \
     python bench.py --steps 8 --warmup 3 --no-e2e $B > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"; tail -3 $OUT/launches.csv | cut -c1-300;;
ncu)
  i=0
  for a in "" "--dtype f64" "--dif-order 0" "--dif-order 0 --dtype f64" "--update-type 3" "--update-type 3 --dif-order 0"; do
    n=$(echo "prof$a" | tr -d ' ' | tr '-' '_')
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:fdtd_update -s 4 -c 1 -f -o $OUT/$n \
       python bench.py $a --steps 6 --warmup 3 --no-e2e $B > $OUT/$n.log 2>&1; echo "ncu [$a] rc=$?"
  done;;
esac
done
ls -la $OUT
