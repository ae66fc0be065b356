cd /root/repo
B=$PWD/openfhe-development_b200
timeout 100 tools/ubench/bfly > gpurun_out/r02g_ubench_occ.txt 2>&1; tail -14 gpurun_out/r02g_ubench_occ.txt | cut -c1-120
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
OFHE_B200_LIB=$B/libofhe_b200_m2.so OFHE_NTT_MODE=split REPS=5 python tools/ntt_prof.py | cut -c1-60,180-300
REPS=5 python tools/ntt_prof.py | cut -c1-60,180-300
timeout 500 python bench.py --steps 20 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -3 gpurun_out/r02g_bench.err; cat gpurun_out/r02g_bench.json
