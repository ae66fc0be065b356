# round-end evidence: GPU tests, bench line, ncu launch list of the bench command, full-set captures of the top kernels
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_final.json; cat gpurun_out/bench_final.json | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json; cat gpurun_out/bench_reference.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r01f_ntt240 python tools/profile_step.py --op ntt240 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r01f_evalmult python tools/profile_step.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
