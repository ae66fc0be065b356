# round-2 session-3 job: tensor-core base conversion -- parity, timing, dependent suites, bench, one ncu capture
cd /root/repo
mkdir -p gpurun_out
export OFHE_BCONV_TC_DEBUG=1
timeout 300 python tools/bconv_tc_check.py parity > gpurun_out/tc_parity.log 2>&1; echo "parity rc=$?"; tail -40 gpurun_out/tc_parity.log
unset OFHE_BCONV_TC_DEBUG
if grep -q "PARITY PASSED" gpurun_out/tc_parity.log; then
  timeout 120 python tools/bconv_tc_check.py time 2>&1 | tail -1 | tee gpurun_out/tc_time_tc1.json
  OFHE_BCONV_TC=0 timeout 120 python tools/bconv_tc_check.py time 2>&1 | tail -1 | tee gpurun_out/tc_time_tc0.json
  timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_bfv.py tests/test_gpu_chain.py tests/test_gpu_shard.py -x -q -m gpu > gpurun_out/tc_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tc_pytest.log
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02p_bench_n1.json; tail -3 gpurun_out/r02p_bench_n1.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:bconv_tc_kernel -c 2 -o gpurun_out/r02p_bconv_tc python tools/bconv_tc_check.py time > gpurun_out/r02p_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02p_ncu.log
fi
