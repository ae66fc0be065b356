cd /root/repo
B=$PWD/openfhe-development_b200
timeout 200 tools/ubench/bfly 2>&1 | grep -E "regcap|error|rror"
OFHE_NTT_MODE=split timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
for lib in "" np; do
  if [ -n "$lib" ]; then export OFHE_B200_LIB=$B/libofhe_b200_$lib.so; else unset OFHE_B200_LIB; fi
  OFHE_NTT_MODE=split OFHE_NTT_CHUNK_MIB=1000 REPS=5 python tools/ntt_prof.py
  OFHE_NTT_MODE=cluster REPS=5 python tools/ntt_prof.py
done
