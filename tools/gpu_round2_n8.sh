#!/bin/bash
# the 8-GPU session of round 2: multi-process tests at world 2/4/8, the bench at N = all GPUs (with the config-4 block and the
# single-GPU invariance check), and BASELINE configurations 3 and 5 at their stated size
OUT=gpurun_out/${1:-r02g}; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu --durations=6 > $OUT/pytest_gpu_multi_${NG}gpu.log 2>&1; echo "multi rc=$?"; tail -4 $OUT/pytest_gpu_multi_${NG}gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 5 \
   > $OUT/bench_n${NG}_auto.json 2> $OUT/bench_n${NG}_auto.err; echo "bench N=$NG rc=$?"; cut -c1-600 $OUT/bench_n${NG}_auto.json; tail -2 $OUT/bench_n${NG}_auto.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $NG --steps 20 --warmup 5 --halo nccl --c4 off --no-invariance \
   > $OUT/bench_n${NG}_nccl.json 2> $OUT/bench_n${NG}_nccl.err; echo "bench nccl N=$NG rc=$?"; cut -c1-300 $OUT/bench_n${NG}_nccl.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 tools/run_configs.py --config c3 c5 \
   2> $OUT/configs.err | grep '^{' > $OUT/configs_c3_c5_n${NG}.jsonl; echo "configs rc=$?"; cut -c1-700 $OUT/configs_c3_c5_n${NG}.jsonl; tail -3 $OUT/configs.err
