#!/bin/bash
OUT=gpurun_out/${1:-r02i}; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 20 --warmup 5 \
   > $OUT/bench_n${NG}_auto.json 2> $OUT/bench_n${NG}_auto.err; echo "bench N=$NG rc=$?"; cut -c1-300 $OUT/bench_n${NG}_auto.json; tail -2 $OUT/bench_n${NG}_auto.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $NG --steps 20 --warmup 5 --halo nccl --c4 off --no-invariance \
   > $OUT/bench_n${NG}_nccl.json 2> $OUT/bench_n${NG}_nccl.err; echo "bench nccl N=$NG rc=$?"; cut -c1-200 $OUT/bench_n${NG}_nccl.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 tools/run_configs.py --config c3 \
   2> $OUT/configs.err | grep '^{' > $OUT/configs_c3_n${NG}.jsonl; echo "configs rc=$?"; cut -c1-300 $OUT/configs_c3_n${NG}.jsonl; tail -2 $OUT/configs.err
