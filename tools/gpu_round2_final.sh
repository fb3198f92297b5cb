#!/bin/bash
# Last GPU-box session of round 2 (after the CTA-order remap left the kernels): the driver's bench line, the reference arm,
# one ncu capture of the fp64 27-point filter kernel, the launch list, smoke, then the whole GPU suite.
# usage (under gpurun): bash tools/gpu_round2_final.sh <tag>
TAG=${1:-r02z}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/nvidia_smi.csv 2>&1
B="--no-cpu-baseline --no-variants --no-like-for-like --c4 off"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$?"; cut -c1-600 $OUT/bench_n1.json
timeout 120 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref_n1.json 2> $OUT/bench_ref_n1.err; cut -c1-300 $OUT/bench_ref_n1.json
timeout 120 bash tools/ncu_job.sh $OUT "--update-type 3 --dtype f64 --dif-order 2" 2>&1 | tail -3
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 8 --warmup 3 --no-e2e $B > $OUT/launches_bench.log 2>&1; echo "launches rc=$?"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
timeout 600 python -m pytest tests -x -q -m gpu --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log; tail -12 $OUT/pytest_gpu.log
