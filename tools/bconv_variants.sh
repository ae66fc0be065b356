for f in "" build/alt6/lib_bc_16.so build/alt6/lib_bc_24.so build/alt6/lib_bc_48.so; do
  if [ -n "$f" ]; then export OFHE_B200_LIB=$PWD/openfhe-development_b200/$f; else unset OFHE_B200_LIB; fi
  python -m pytest tests/test_gpu_parity.py tests/test_bfv.py -m gpu -x -q 2>&1 | tail -1
  python tools/bench_kernels.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', 'bconv', round(d['bconv_15to45_us'],1), 'modup', round(d['modup_us'],1), 'moddown', round(d['moddown2_us'],1), 'evalmult', round(d['evalmult_us'],1))"
done
