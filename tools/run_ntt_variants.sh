#!/bin/bash
cd /root/repo
B=$PWD/openfhe-development_b200
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I$B/csrc -o tools/ubench/bfly tools/ubench/bfly.cu 2>/dev/null
timeout 120 tools/ubench/bfly > gpurun_out/r02k_ubench_bfly.txt 2>&1
p() { python tools/ntt_prof.py | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-22s %-8s stream=%s L=%s logN=%s fwd %.3f inv %.3f' % (d['lib'], d['mode'], '$OFHE_NTT_STREAM', d['limbs'], d['logN'], d['fwd_us_per_limb'], d['inv_us_per_limb']))"; }
export OFHE_NTT_MODE=split REPS=5 OFHE_NTT_STREAM=0
for v in default fdf0 idf0 x3 x4 x5 x6; do
  L=$B/libofhe_b200.so; [ $v != default ] && L=$B/libofhe_b200_$v.so
  OFHE_B200_LIB=$L p
done 2>&1 | tee gpurun_out/r02k_ntt_variants.txt
