# round-end evidence: GPU tests, bench lines (ours + reference arm), warm per-kernel launch list of one EvalMult
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_final.json; cut -c1-300 gpurun_out/bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json; cut -c1-200 gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --profile-from-start off --csv --log-file gpurun_out/r01g_launches_warm.csv python tools/profile_step.py > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
