# ModDown tail fused into the forward NTT's tile pass: parity suites (incl. the full config-2 and N=2^17 sizes), kernel timings, launch list, bench with aux
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_bfv.py tests/test_gpu_chain.py tests/test_gpu_shard.py -x -q -m gpu > gpurun_out/tail_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/tail_pytest.log
timeout 300 python tools/bench_kernels.py 2>&1 | tail -1 | tee gpurun_out/r02u_kernels.json
OFHE_NTT_TAIL=0 timeout 300 python tools/bench_kernels.py 2>&1 | tail -1 | tee gpurun_out/r02u_kernels_notail.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02u_launches_evalmult.csv python tools/profile_step.py > gpurun_out/r02u_prof.log 2>&1; echo "launch list rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02u_launches_evalmult.csv')) if len(r)>10 and r[0].isdigit()]
tot=0
for r in rows:
    name=r[4].split('(')[0].split('::')[-1][:40]; t=float(r[-1])/1e3; tot+=t
    print(f"{name:42s} grid {r[8]:16s} {t:8.1f} us")
print("sum", round(tot,1))
PY
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02u_bench_n1.json 2> gpurun_out/r02u_bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02u_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','gpu_launches')})
print(json.dumps(d.get('roofline'))[:300])
a=d.get('aux',{})
for k,v in a.items():
    print(k, json.dumps(v)[:400])
PY
tail -3 gpurun_out/r02u_bench_n1.err
