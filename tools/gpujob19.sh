# four 128-column accumulator buffers instead of two 256-column ones (OFHE_TC_NBUF=4): parity + timing against the default
cd /root/repo
mkdir -p gpurun_out
V=$PWD/openfhe-development_b200/libofhe_b200_nb4.so
OFHE_BCONV_TC_DEBUG=1 OFHE_B200_LIB=$V timeout 200 python tools/bconv_tc_check.py parity > gpurun_out/tc5_parity_nb4.log 2>&1; echo "parity(nb4) rc=$?"; tail -9 gpurun_out/tc5_parity_nb4.log
OFHE_B200_LIB=$V timeout 100 python tools/bconv_tc_check.py time 2>&1 | tail -1 | tee gpurun_out/tc5_time_nb4.json
timeout 100 python tools/bconv_tc_check.py time 2>&1 | tail -1 | tee gpurun_out/tc5_time_nb2.json
OFHE_B200_LIB=$V timeout 200 python tools/bench_kernels.py 2>&1 | tail -1 | tee gpurun_out/r02z_kernels_nb4.json
