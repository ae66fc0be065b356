# last checks of the round: the remaining reference suites on the device path (automorphism maps now built on a private
# stream), in-flight sweep of the single-GPU bench line
cd /root/repo
mkdir -p gpurun_out
export OFHE_B200_HAL_STATS=1
timeout 420 python -m pytest tests/test_gpu_hal.py -x -q -m gpu -k "automorphism or general_she or bootstrap_driver" --durations=4 > gpurun_out/r02y_hal_rest.log 2>&1; echo "hal rest rc=$?"; tail -8 gpurun_out/r02y_hal_rest.log
for S in 8 6; do
  timeout 240 python bench.py --steps 24 --warmup 3 --slots $S --no-cpu-baseline --no-aux > gpurun_out/r02y_bench_s$S.json 2> gpurun_out/r02y_bench_s$S.err; echo "slots=$S rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/r02y_bench_s$S.json').read().strip().splitlines()[-1]); print({k:d.get(k) for k in ('value','ms_per_step')}, d['e2e']['value'])"
done
