"""profiles/<summary>.csv (tools/ncu_summarize.py) -> profiles/ntt_roofline_traffic.json: DRAM bytes of ONE forward
limb-NTT launch over the roofline slab (all kernels of the launch summed), which bench.py reports as roofline.traffic.
usage: python tools/ncu_traffic.py profiles/r02e_split_summary.csv ntt_fwd 240 65536"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, pat, limbs, N = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(open(path)))
hdr = [h.split("[")[0] for h in rows[0]]
units = {h.split("[")[0]: (h.split("[")[1].rstrip("]") if "[" in h else "") for h in rows[0]}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
seen, total, kernels = set(), 0.0, []
for r in rows[1:]:
    d = dict(zip(hdr, r))
    if pat not in d["kernel"] or d["kernel"] in seen:
        continue
    seen.add(d["kernel"])  # one instance of each kernel of the launch
    b = float(d["dram_read"]) * scale[units["dram_read"]] + float(d["dram_write"]) * scale[units["dram_write"]]
    total += b
    kernels.append({"kernel": d["kernel"], "dram_bytes": b, "duration_us": float(d["duration_us"])})
out = {"dram_bytes_per_launch": total, "limbs": limbs, "N": N, "algorithmic_bytes": 16.0 * N * limbs,
       "ratio_to_algorithmic": total / (16.0 * N * limbs), "kernels": kernels,
       "source": f"{os.path.relpath(path, ROOT)} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "ntt_roofline_traffic.json"), "w"), indent=1)
print(json.dumps(out))
