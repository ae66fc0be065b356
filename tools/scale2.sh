set -x
for mode in "--slots 2" "--slots 3" "--slots 4"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 24 --warmup 4 $mode 2>&1 | tail -1 | cut -c1-230
done
