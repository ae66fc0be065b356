cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_hal.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_hal.py -x -q -k "evalmult or bgv or automorphism" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02m_bench_n1.json 2> gpurun_out/r02m_bench_n1.err; tail -2 gpurun_out/r02m_bench_n1.err; cut -c1-700 gpurun_out/r02m_bench_n1.json
OFHE_NTT_MODE=split REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ntt_(fwd|inv)_(tile|strided)' -c 4 -o gpurun_out/r02m_ntt python tools/ntt_prof.py > gpurun_out/r02m_ntt.log 2>&1; tail -1 gpurun_out/r02m_ntt.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02m_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02m_bench_under_ncu.log 2>&1; tail -1 gpurun_out/r02m_bench_under_ncu.log | cut -c1-200
