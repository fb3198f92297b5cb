#!/bin/bash
# Round-2 GPU-box session. usage (under gpurun): bash tools/gpu_round2.sh <tag> [what...]
#   what in {tests, multi, smoke, bench, ref, launches, ncu, benchN}
TAG=${1:-r02}; shift
WHAT=${@:-tests smoke bench ref launches ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/nvidia_smi.csv 2>&1
nproc > $OUT/nproc.txt; free -g | head -2 >> $OUT/nproc.txt
NG=$(nvidia-smi -L | wc -l)
B="--no-cpu-baseline --no-variants --no-like-for-like --c4 off"
for w in $WHAT; do
case $w in
tests)
  timeout 2400 python -m pytest tests -x -q -m gpu --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -25 $OUT/pytest_gpu.log;;
multi)
  timeout 1500 python -m pytest tests/test_gpu_multi.py -x -q -m gpu --durations=10 > $OUT/pytest_gpu_multi_${NG}gpu.log 2>&1; echo "multi rc=$?" | tee -a $OUT/pytest_gpu_multi_${NG}gpu.log; tail -15 $OUT/pytest_gpu_multi_${NG}gpu.log;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log;;
bench)   # the driver's command line
  timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; cut -c1-2500 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err;;
ref)
  timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref_n1.json 2> $OUT/bench_ref_n1.err; cut -c1-1200 $OUT/bench_ref_n1.json
  timeout 600 python bench.py --impl reference --workload c1 --steps 500 --warmup 5 > $OUT/bench_ref_c1.json 2> $OUT/bench_ref_c1.err; cut -c1-400 $OUT/bench_ref_c1.json;;
benchN)  # one process per GPU on every GPU of the box, both halo transports
  for h in auto nccl; do
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 5 --halo $h \
      > $OUT/bench_n${NG}_$h.json 2> $OUT/bench_n${NG}_$h.err; echo "bench N=$NG halo=$h rc=$?"; cut -c1-3000 $OUT/bench_n${NG}_$h.json; tail -3 $OUT/bench_n${NG}_$h.err
  done;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
     python bench.py --steps 8 --warmup 3 --no-e2e $B > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"; tail -3 $OUT/launches.csv | cut -c1-300;;
ncu)
  for a in "" "--update-type 2" "--update-type 3 --dtype f64" "--update-type 3"; do
    n=$(echo "prof$a" | tr -d ' ' | tr '-' '_')
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:fdtd_update -s 4 -c 1 -f -o $OUT/$n \
       python bench.py $a --steps 6 --warmup 3 --no-e2e $B > $OUT/$n.log 2>&1; echo "ncu [$a] rc=$?"
  done;;
esac
done
ls -la $OUT
