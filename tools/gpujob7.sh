cd /root/repo
export OFHE_B200_HAL_STATS=1
for f in '*UTCKKSRNS_SER*' '*UTBGVRNS_SER*' 'UTBFVRNS_SER*' '*BOOTSTRAP_SERIALIZE*'; do
  n=$(echo $f | tr -dc 'A-Z_'); timeout 600 oracle/_ref/hal/pke_tests_b200 --gtest_filter="$f" > gpurun_out/hal3_$n.log 2>&1
  echo "== $n rc=$? $(grep -E 'tests ran|PASSED|FAILED TEST|hal\] device' gpurun_out/hal3_$n.log | tr '\n' ' ' | cut -c1-260)"
done
nproc; free -g | head -2
B=oracle/_ref/hal
(timeout 300 $B/bootstrap_b200 --logn 14 --slots 0 2> gpurun_out/boot_b200_14.err || echo "b200 n14 rc=$?") | tee gpurun_out/boot_b200_14.json | cut -c1-500
(timeout 600 $B/bootstrap_b200 --logn 16 --slots 0 2> gpurun_out/boot_b200_16.err || echo "b200 n16 rc=$?") | tee gpurun_out/boot_b200_16.json | cut -c1-500
(timeout 900 $B/bootstrap_b200 --logn 17 --slots 0 2> gpurun_out/boot_b200_17.err || echo "b200 n17 rc=$?") | tee gpurun_out/boot_b200_17.json | cut -c1-500
tail -2 gpurun_out/boot_b200_17.err
(OMP_NUM_THREADS=32 OMP_PROC_BIND=spread OMP_PLACES=cores timeout 900 $B/bootstrap_ref --logn 17 --slots 0 --warm 0 2> gpurun_out/boot_ref_17.err || echo "ref n17 rc=$?") | tee gpurun_out/boot_ref_17.json | cut -c1-500
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02n_bench_aux.json 2> gpurun_out/r02n_bench_aux.err; tail -1 gpurun_out/r02n_bench_aux.err; python -c "
import json; d=json.load(open('gpurun_out/r02n_bench_aux.json')); print(json.dumps(d['aux']['config2_resident_chain_and_rotations'])[:900])"
