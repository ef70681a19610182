N=${1:-8}
for t in 37 74 148 296 592; do
  echo "tile $t: $(PQA_B200_PIPE_TILE=$t ./probqa_b200/lib/pqa_shard_launcher --gpus $N --axis targets --exact-order --questions 10000 --answers 5 --targets 100000 --batch 64 --steps 10 --warmup 3 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['phases_ms'])")"
done
