#!/bin/bash
# round-2 measurement of the streaming tile kernels: parity first, then timings of the variants, then one ncu capture
cd /root/repo
B=$PWD/openfhe-development_b200
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "ntt or kat or config2 or config4" 2>&1 | tail -3
for v in default w3 w5; do
  L=$B/libofhe_b200.so; [ $v != default ] && L=$B/libofhe_b200_$v.so
  OFHE_B200_LIB=$L OFHE_NTT_MODE=split REPS=5 timeout 120 python tools/ntt_prof.py | cut -c1-40,170-300
done
OFHE_NTT_STREAM=0 OFHE_NTT_MODE=split REPS=5 timeout 120 python tools/ntt_prof.py | cut -c1-40,170-300
REPS=5 timeout 120 python tools/ntt_prof.py | cut -c1-40,170-300
LIMBS=45 OFHE_NTT_MODE=split REPS=5 timeout 120 python tools/ntt_prof.py | cut -c1-40,170-300
LIMBS=45 OFHE_NTT_STREAM=0 OFHE_NTT_MODE=split REPS=5 timeout 120 python tools/ntt_prof.py | cut -c1-40,170-300
LOGN=15 OFHE_NTT_MODE=split REPS=5 timeout 120 python tools/ntt_prof.py | cut -c1-40,170-300
LOGN=17 LIMBS=120 OFHE_NTT_MODE=split REPS=5 timeout 120 python tools/ntt_prof.py | cut -c1-40,170-300
OFHE_NTT_MODE=split REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ntt_(fwd|inv)_tile_stream' -c 2 -o gpurun_out/r02i_stream python tools/ntt_prof.py > gpurun_out/r02i_stream.log 2>&1
tail -2 gpurun_out/r02i_stream.log
