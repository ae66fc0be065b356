# scaling run on an 8-GPU box: sharded parity (NCCL + peer-push) at world 8 / 4, then bench.py at 8 and 4
set -x
for W in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29517 tools/shard_check.py 2>&1 | grep -E "OK|MISMATCH|Error|error|Traceback" | sort | uniq -c | head -20
done
for mode in "--gpus 8 --exchange push" "--gpus 8 --exchange push --no-graph" "--gpus 4 --exchange push"; do
  W=$(echo $mode | cut -d' ' -f2)
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29518 bench.py --steps 20 --warmup 3 $mode 2>&1 | tail -2 | cut -c1-900 | tee -a gpurun_out/scale_push.jsonl
done
