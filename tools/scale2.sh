set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29517 tools/shard_check.py 2>&1 | grep -E "OK|MISMATCH|Error|error|Traceback" | head -12
for mode in "--exchange push" "--exchange push --no-graph" "--exchange nccl"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 $mode 2>&1 | tail -2 | cut -c1-700
done
