# tensor-core base conversion v2 (decoupled warps, pipelined TMEM loads, single-word reduction): parity, timing, suites, bench, ncu
cd /root/repo
mkdir -p gpurun_out
export OFHE_BCONV_TC_DEBUG=1
timeout 300 python tools/bconv_tc_check.py parity > gpurun_out/tc2_parity.log 2>&1; echo "parity rc=$?"; tail -30 gpurun_out/tc2_parity.log
OFHE_B200_LIB=$PWD/openfhe-development_b200/libofhe_b200_slowred.so timeout 300 python tools/bconv_tc_check.py parity > gpurun_out/tc2_parity_slowred.log 2>&1; echo "parity(slowred) rc=$?"; tail -3 gpurun_out/tc2_parity_slowred.log
unset OFHE_BCONV_TC_DEBUG
timeout 120 python tools/bconv_tc_check.py time 2>&1 | tail -1 | tee gpurun_out/tc2_time_fast.json
OFHE_B200_LIB=$PWD/openfhe-development_b200/libofhe_b200_slowred.so timeout 120 python tools/bconv_tc_check.py time 2>&1 | tail -1 | tee gpurun_out/tc2_time_slowred.json
if grep -q "PARITY PASSED" gpurun_out/tc2_parity.log; then
  timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_bfv.py tests/test_gpu_chain.py tests/test_gpu_shard.py -x -q -m gpu > gpurun_out/tc2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tc2_pytest.log
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_bench_n1.json 2> gpurun_out/r02q_bench_n1.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02q_bench_n1.json; tail -3 gpurun_out/r02q_bench_n1.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:bconv_tc_kernel -c 2 -o gpurun_out/r02q_bconv_tc python tools/bconv_tc_check.py time > gpurun_out/r02q_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02q_ncu.log
fi
