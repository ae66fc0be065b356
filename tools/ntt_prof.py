"""Forward / inverse limb-NTT on the roofline slab (240 limbs of N = 2^16, 4 limbs per modulus, modulus-major order)
for profiling:   ncu --set full ... python tools/ntt_prof.py   (REPS launches of each direction)
Prints CUDA-event timings when run bare."""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
import ofhe_b200 as m  # noqa: E402

logn = int(os.environ.get("LOGN", "16"))
L = int(os.environ.get("LIMBS", "240"))
reps = int(os.environ.get("REPS", "3"))
p = m.ckks_paramgen(1 << logn, 44, 59, 60, 0, 0)
N, nmod = p["N"], len(p["moduli"])
ctx = m.Context(0)
ring = m.Ring(ctx, N, p["moduli"], p["roots"])
occ = (C.c_int * 2)()
m.lib().ofhe_ntt_cluster_occupancy(ctx.h, logn, occ)
slab = ctx.zeros(L, N)
per = max(1, L // nmod)
lidx = [min(i // per, nmod - 1) for i in range(L)]  # modulus-major: limbs sharing a modulus are adjacent
ring.fill_uniform(slab, lidx, 3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = {"lib": os.path.basename(os.environ.get("OFHE_B200_LIB", "default")), "mode": os.environ.get("OFHE_NTT_MODE", "cluster"), "chunk_mib": os.environ.get("OFHE_NTT_CHUNK_MIB", "32"), "logN": logn, "limbs": L, "max_active_clusters_fwd_inv": [occ[0], occ[1]], "sm_count": ctx.sm_count}
for name, fn in (("fwd", ring.ntt_fwd), ("inv", ring.ntt_inv)):
    fn(slab, slab, lidx)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn(slab, slab, lidx)
    e1.record()
    torch.cuda.synchronize()
    out[f"{name}_us_per_limb"] = e0.elapsed_time(e1) / reps * 1e3 / L
print(json.dumps(out))
