# compare NTT builds: default lib + each build/alt2/lib_ntt_*.so (parity subset first, then kernel timings)
for f in "" build/alt2/lib_ntt_base.so build/alt2/lib_ntt_asm.so build/alt2/lib_ntt_noskip.so; do
  if [ -n "$f" ]; then export OFHE_B200_LIB=$PWD/openfhe-development_b200/$f; else unset OFHE_B200_LIB; fi
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt or golden or kat or evalmult" 2>&1 | tail -1
  python tools/bench_kernels.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', {k: round(v,2) for k,v in d.items() if isinstance(v,float)})"
done
