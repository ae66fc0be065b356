# round-end style scaling run on an N-GPU box: sharded-parity check at world 8/4, then bench at 8, 4, 2 (and 1)
set -x
NG=$(python -c "import torch; print(torch.cuda.device_count())")
for W in 8 4; do
  [ $W -le $NG ] && timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29517 tools/shard_check.py 2>&1 | grep -E "OK|MISMATCH|Error|error" | head -12
done
for W in 8 4 2; do
  [ $W -le $NG ] && timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$W --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $W --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/scale_$W.json
done
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/scale_1.json
