#!/bin/bash
# the reference's own gtest suites, relinked against the B200 lattice backend (oracle/Makefile.hal), on the GPU box
cd /root/repo
H=oracle/_ref/hal
export OFHE_B200_HAL_STATS=1
run() { # name binary filter limit
  local t0=$(date +%s)
  timeout $4 $H/$2 --gtest_filter="$3" > gpurun_out/hal_$1.log 2>&1
  echo "== $1 rc=$? $(( $(date +%s) - t0 ))s: $(grep -E 'tests ran|PASSED|FAILED TEST|hal\] device' gpurun_out/hal_$1.log | tr '\n' ' ' | cut -c1-400)"
}
run core core_tests_b200 'UTNTT.*:UTTransform.*:UTDCRTPoly.*' 300
run bfvcrt pke_tests_b200 'UTBFVRNS_CRT.*' 300
run evalmult pke_tests_b200 '*UTGENERAL_EVALMULT*' 400
run ckks pke_tests_b200 '*UTCKKSRNS.*' 600
run bgv pke_tests_b200 '*UTBGVRNS.*' 400
run bfv pke_tests_b200 '*UTBFVRNS.*' 400
run ckksauto pke_tests_b200 '*UTCKKSRNS_AUTOMORPHISM*' 300
run she pke_tests_b200 '*UTGENERAL_SHE*' 400
run boot pke_tests_b200 "*UTCKKSRNS_BOOT*" 900
