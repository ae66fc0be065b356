cd /root/repo
timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_hal.py 2>&1 | tail -4
export OFHE_B200_HAL_STATS=1
for f in '*UTCKKSRNS_BOOT*' '*UTCKKSRNS_SER*'; do
  n=$(echo $f | tr -dc 'A-Z_'); timeout 900 oracle/_ref/hal/pke_tests_b200 --gtest_filter="$f" > gpurun_out/hal2_$n.log 2>&1
  echo "== $n rc=$? $(grep -E 'tests ran|PASSED|FAILED TEST|hal\] device' gpurun_out/hal2_$n.log | tr '\n' ' ' | cut -c1-300)"
done
timeout 900 python -m pytest tests/test_gpu_hal.py -x -q -k "evalmult or bgv or automorphism or ckksrns and not boot and not serial" 2>&1 | tail -3
timeout 700 python bench.py --steps 20 --warmup 3 > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err; tail -2 gpurun_out/r02m_bench_n1.err; cut -c1-600 gpurun_out/r02m_bench_n1.json; echo
OFHE_NTT_MODE=split REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ntt_(fwd|inv)_(tile|strided)' -c 4 -o gpurun_out/r02m_ntt python tools/ntt_prof.py > gpurun_out/r02m_ntt.log 2>&1; tail -1 gpurun_out/r02m_ntt.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02m_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02m_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r02m_bench_under_ncu.log | cut -c1-120
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "kat or golden or limb_ops or approx_switch or upload_cereal" > gpurun_out/r02m_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02m_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "kat or ntt_golden or limb_ops" > gpurun_out/r02m_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02m_sanitizer_racecheck.log
