# 4-GPU limb-sharded bench line: pipeline-slot sweep (latency-bound regime)
cd /root/repo
mkdir -p gpurun_out
for S in 4 8 12; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 24 --warmup 3 --slots $S --no-cpu-baseline --no-aux > gpurun_out/r02v_bench_n4_s$S.json 2> gpurun_out/r02v_bench_n4_s$S.err
  echo "n4 slots=$S rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02v_bench_n4_s$S.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step')}, d.get('replicas',{}).get('value'), d.get('e2e',{}).get('value'))
except Exception as e:
    print('no json',e)
PY
  tail -2 gpurun_out/r02v_bench_n4_s$S.err | cut -c1-300
done
