# sharded rotation / rescale on virtual ranks (CUDA path), the fused-tail variant test, smoke()
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_shard.py -x -q -m gpu --durations=5 > gpurun_out/r02x_shard_pytest.log 2>&1; echo "shard pytest rc=$?"; tail -12 gpurun_out/r02x_shard_pytest.log
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_parity.py -x -q -m gpu -k "chain or bsgs or config4" > gpurun_out/r02x_tail_pytest.log 2>&1; echo "rotation users rc=$?"; tail -3 gpurun_out/r02x_tail_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
