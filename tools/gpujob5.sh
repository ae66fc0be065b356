cd /root/repo
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02l_bench_n1.json 2> gpurun_out/r02l_bench_n1.err; tail -2 gpurun_out/r02l_bench_n1.err; cut -c1-900 gpurun_out/r02l_bench_n1.json
