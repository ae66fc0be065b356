for f in "" build/alt/lib_bc_8_8.so build/alt/lib_bc_16_1.so build/alt/lib_bc_16_4.so; do
  if [ -n "$f" ]; then export OFHE_B200_LIB=$PWD/openfhe-development_b200/$f; else unset OFHE_B200_LIB; fi
  python tools/bench_kernels.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', 'bconv', round(d['bconv_15to45_us'],1), 'modup', round(d['modup_us'],1), 'moddown', round(d['moddown2_us'],1), 'evalmult', round(d['evalmult_us'],1))"
done
