cd /root/repo
bash tools/run_ntt_variants.sh
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "upload_cereal or approx_switch" 2>&1 | tail -3
bash tools/run_hal.sh
