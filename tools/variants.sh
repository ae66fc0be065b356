# usage: tools/variants.sh lib1.so lib2.so ...   (paths relative to openfhe-development_b200/; "" = the default build)
for f in "" "$@"; do
  if [ -n "$f" ]; then export OFHE_B200_LIB=$PWD/openfhe-development_b200/$f; else unset OFHE_B200_LIB; fi
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  python tools/bench_kernels.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', {k: round(v,1) for k,v in d.items() if isinstance(v,float)})"
done
