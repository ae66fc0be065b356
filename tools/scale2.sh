set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29517 tools/shard_check.py 2>&1 | grep -E "OK|MISMATCH|Error|error|Traceback" | sort | uniq -c | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/scale2_final.json | cut -c1-260
