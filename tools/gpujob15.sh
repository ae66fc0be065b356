# the reference's own unit tests on the device path with the tensor-core base conversion in the backend + bench line with aux
cd /root/repo
mkdir -p gpurun_out
export OFHE_B200_HAL_STATS=1
timeout 700 python -m pytest tests/test_gpu_hal.py -x -q -m gpu -k "core_ntt or bfvrns_crt or general_evalmult or test_ckksrns or bgvrns" --durations=8 > gpurun_out/r02w_hal_pytest.log 2>&1; echo "hal pytest rc=$?"; tail -14 gpurun_out/r02w_hal_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02w_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches')})
print(json.dumps(d.get('aux',{}).get('base_conversion_15_45'))[:400])
PY
tail -2 gpurun_out/r02w_bench_n1.err
