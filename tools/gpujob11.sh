# tensor-core base conversion v3: MMA rates by kind, parity, timing (stand-alone and in the key switch), launch list, suites, bench
cd /root/repo
mkdir -p gpurun_out
export OFHE_BCONV_TC_DEBUG=1
timeout 300 python tools/bconv_tc_check.py parity > gpurun_out/tc4_parity.log 2>&1; echo "parity rc=$?"; tail -12 gpurun_out/tc4_parity.log
OFHE_B200_LIB=$PWD/openfhe-development_b200/libofhe_b200_db.so timeout 300 python tools/bconv_tc_check.py parity > gpurun_out/tc4_parity_db.log 2>&1; echo "parity(db) rc=$?"; tail -2 gpurun_out/tc4_parity_db.log
unset OFHE_BCONV_TC_DEBUG
timeout 120 python tools/bconv_tc_check.py time 2>&1 | tail -1 | tee gpurun_out/tc4_time.json
OFHE_B200_LIB=$PWD/openfhe-development_b200/libofhe_b200_db.so timeout 120 python tools/bconv_tc_check.py time 2>&1 | tail -1 | tee gpurun_out/tc4_time_db.json
if grep -q "PARITY PASSED" gpurun_out/tc4_parity.log; then
  timeout 300 python tools/bench_kernels.py 2>&1 | tail -1 | tee gpurun_out/r02s_kernels_tc.json
  OFHE_B200_LIB=$PWD/openfhe-development_b200/libofhe_b200_db.so timeout 300 python tools/bench_kernels.py 2>&1 | tail -1 | tee gpurun_out/r02s_kernels_tc_db.json
  OFHE_BCONV_TC=0 timeout 300 python tools/bench_kernels.py 2>&1 | tail -1 | tee gpurun_out/r02s_kernels_cuda.json
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02s_launches_evalmult.csv python tools/profile_step.py > gpurun_out/r02s_prof.log 2>&1; echo "launch list rc=$?"
  python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02s_launches_evalmult.csv')) if len(r)>10 and r[0].isdigit()]
tot=0
for r in rows:
    name=r[4].split('(')[0].split('::')[-1][:40]; t=float(r[-1])/1e3; tot+=t
    print(f"{name:42s} grid {r[8]:16s} {t:8.1f} us")
print("sum", round(tot,1))
PY
  timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_bfv.py tests/test_gpu_chain.py tests/test_gpu_shard.py -x -q -m gpu > gpurun_out/tc4_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tc4_pytest.log
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r02s_bench_n1.json; tail -3 gpurun_out/r02s_bench_n1.err
fi
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bconv_tc_kernel -c 2 -o gpurun_out/r02s_bconv_tc_pipeline python tools/profile_step.py > gpurun_out/r02s_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02s_ncu.log
