#!/bin/bash
# Multi-GPU checks and benches of the sharded modes (run under `gpurun --gpus N -- bash scripts/run_multi_gpu.sh N [full|config4]`).
# Output: gpurun_out/multi_gpu_N.log
N=${1:-2}
MODE=${2:-full}
mkdir -p gpurun_out
LOG=gpurun_out/multi_gpu_${N}.log
: > $LOG
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
PORT=29510
run() { PORT=$((PORT+1)); echo "### $*" >> $LOG; timeout 600 "$@" >> $LOG 2>&1; echo "### rc=$?" >> $LOG; }
nvidia-smi topo -m 2>&1 | head -12 >> $LOG
run $TR --master-port $PORT scripts/p2p_multiproc_check.py
if [ "$MODE" = "exact8" ]; then
  run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange p2p --workload 10000x5x100000_b64 --steps 5 --warmup 3
  run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange p2p --exact-order --workload 10000x5x100000_b64 --steps 5 --warmup 3
  grep -v "^\[W\|^W0\|^\*\*\*\|Setting OMP\|^$\|NCCL version" $LOG | cut -c1-420 | tail -30
  exit 0
fi
if [ "$MODE" = "exact" ]; then
  run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange p2p --steps 200 --warmup 20
  run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange p2p --exact-order --steps 200 --warmup 20
  run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange p2p --workload 10000x5x100000_b64 --steps 5 --warmup 3
  run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange p2p --exact-order --workload 10000x5x100000_b64 --steps 5 --warmup 3
  grep -v "^\[W\|^W0\|^\*\*\*\|Setting OMP\|^$\|NCCL version" $LOG | cut -c1-420 | tail -30
  exit 0
fi
if [ "$MODE" = "full" ]; then
  for X in p2p nccl; do
    run $TR --master-port $PORT bench.py --gpus $N --shard questions --exchange $X --steps 200 --warmup 20
    run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange $X --steps 200 --warmup 20
    run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange $X --workload 2000x5x20000_b64 --steps 20 --warmup 3
  done
fi
# BASELINE config 4: 10000 x 5 x 100000 (48 GB KB) over the N GPUs, and the same workload on one GPU for the scaling ratio
run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange p2p --workload 10000x5x100000_b64 --steps 5 --warmup 3
run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange p2p --exact-order --workload 10000x5x100000_b64 --steps 5 --warmup 3
run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange nccl --workload 10000x5x100000_b64 --steps 5 --warmup 3
run $TR --master-port $PORT bench.py --gpus $N --shard questions --exchange p2p --workload 10000x5x100000_b64 --steps 5 --warmup 3
run $TR --master-port $PORT bench.py --gpus $N --shard targets --exchange p2p --workload 10000x5x100000_b256 --steps 3 --warmup 3
run python bench.py --gpus 1 --workload 10000x5x100000_b64 --steps 3 --warmup 3 --no-cpu-baseline
run $TR --master-port $PORT bench.py --gpus $N --steps 100 --warmup 10
grep -v "^\[W\|^W0\|^\*\*\*\|Setting OMP\|^$\|NCCL version" $LOG | cut -c1-420 | tail -60
