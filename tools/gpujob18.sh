# immortal HAL runtime (static-destruction-order fix): the bootstrap driver and two gtest suites on the device path
cd /root/repo
mkdir -p gpurun_out
export OFHE_B200_HAL_STATS=1
timeout 420 python -m pytest tests/test_gpu_hal.py -x -q -m gpu -k "bootstrap_driver or core_ntt or general_evalmult or automorphism" --durations=4 > gpurun_out/r02z_hal.log 2>&1; echo "hal rc=$?"; tail -8 gpurun_out/r02z_hal.log
