#!/bin/bash
# Round-2 multi-GPU evidence on N GPUs of one box (gpurun --gpus N): the driver's own bench command (replicas + the embedded
# BASELINE config 4 target-sharded leg), the C++ one-process-per-GPU launcher, the in-process ShardGroup, and the tests that
# need more than one GPU. Everything lands in gpurun_out/r02_multi_gpu_$N.log.
N=${1:-2}
OUT=gpurun_out/r02_multi_gpu_$N.log
mkdir -p gpurun_out; : > $OUT
run() { echo "### $*" >> $OUT; "$@" >> $OUT 2>> $OUT.err; echo >> $OUT; }
nvidia-smi topo -m 2>/dev/null | head -12 >> $OUT
run python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 50 --warmup 10
run ./probqa_b200/lib/pqa_shard_launcher --gpus $N --axis targets --exact-order --questions 10000 --answers 5 --targets 100000 --batch 64 --steps 10 --warmup 3
run ./probqa_b200/lib/pqa_shard_launcher --gpus $N --axis targets --questions 10000 --answers 5 --targets 100000 --batch 64 --steps 10 --warmup 3
run ./probqa_b200/lib/pqa_shard_launcher --gpus $N --axis questions --questions 10000 --answers 5 --targets 100000 --batch 64 --steps 10 --warmup 3
run python bench.py --group $N --shard targets --exact-order --workload 10000x5x100000_b64 --steps 10 --warmup 3
run python bench.py --group $N --shard targets --workload 10000x5x100000_b64 --steps 10 --warmup 3
run python -m pytest tests/test_gpu_group.py tests/test_gpu_sharded.py tests/test_gpu_client.py -q -m gpu -k "group or across_processes or launcher"
tail -c 1500 $OUT.err >> $OUT
grep -c . $OUT
