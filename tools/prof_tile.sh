cd /root/repo
export OFHE_NTT_MODE=split REPS=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'ntt_(fwd|inv)_tile' -c 4 -o gpurun_out/r02h_tile python tools/ntt_prof.py > gpurun_out/r02h_tile.log 2>&1
tail -3 gpurun_out/r02h_tile.log
REPS=5 timeout 120 python tools/ntt_prof.py
