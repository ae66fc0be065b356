# 2-GPU check of the limb-sharded bench line with the tensor-core base conversion (parity gate in-run) + pipeline-slot sweep
cd /root/repo
mkdir -p gpurun_out
for S in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --slots $S --no-cpu-baseline --no-aux > gpurun_out/r02t_bench_n2_s$S.json 2> gpurun_out/r02t_bench_n2_s$S.err
  echo "n2 slots=$S rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02t_bench_n2_s$S.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','parity_checked','replicas')})
except Exception as e:
    print('no json',e)
PY
  tail -2 gpurun_out/r02t_bench_n2_s$S.err
done
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
