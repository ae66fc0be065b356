"""BASELINE config 3: BFV EvalMult (HPSPOVERQ) at N = 2^15, 12 + 12 limbs (+ 4 key-switch limbs, dnum 3) on one GPU,
with the unmodified reference timed beside it on the host cores when oracle/_ref is present.
Tables: tests/golden/bfv_cfg3_tables.npz (the reference's own CryptoParametersBFVRNS tables)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ofhe_b200 as m  # noqa: E402
from oracle_lib import Reference, reference_available  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "bfv_cfg3_tables.npz"))
N, nQ, nR = (int(v) for v in g["shape"])
sizeP, dnum, _ = (int(v) for v in g["ks_shape"])
T = {k[2:]: g[k] for k in g.files if k.startswith("T_")}
ctx = m.Context(0)
ring = m.Ring(ctx, N, g["moduli"], g["roots"])
mult = m.BfvMult(ring, nQ, nR, T)
ring_ks = m.Ring(ctx, N, np.concatenate([g["moduli"][:nQ], g["moduliP"]]), np.concatenate([g["roots"][:nQ], g["rootsP"]]))
hks = m.Hks(ring_ks, nQ, sizeP, dnum)
evk = m.EvalKey(hks, seed=3)
ct1, ct2 = ctx.zeros(2, nQ, N), ctx.zeros(2, nQ, N)
ring.fill_uniform(ct1.view(2 * nQ, N), list(range(nQ)) * 2, 1)
ring.fill_uniform(ct2.view(2 * nQ, N), list(range(nQ)) * 2, 2)
out3, out = ctx.zeros(3, nQ, N), ctx.zeros(2, nQ, N)


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


res = {"config": "BFV EvalMult HPSPOVERQ, N=2^15, 12 Q + 12 R limbs (60-bit), key switch HYBRID 4 P limbs dnum 3",
       "norelin_us": timeit(lambda: mult.evalmult_norelin(ct1, ct2, out3)),
       "evalmult_relin_us": timeit(lambda: mult.evalmult(hks, ct1, ct2, evk, out))}
res["evalmult_relin_per_s"] = 1e6 / res["evalmult_relin_us"]
if reference_available():
    b = Reference().bfv(N, 20, 60, 65537, 0)
    Q = b.moduli[:b.sizeQ]
    rng = np.random.default_rng(5)
    c1 = np.stack([rng.integers(0, q, size=b.N, dtype=np.uint64) for q in np.tile(Q, 2)]).reshape(2, b.sizeQ, b.N)
    c2 = np.stack([rng.integers(0, q, size=b.N, dtype=np.uint64) for q in np.tile(Q, 2)]).reshape(2, b.sizeQ, b.N)
    b.evalmult(c1, c2)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        b.evalmult(c1, c2)
    res["reference_cpu_ms"] = (time.perf_counter() - t0) / reps * 1e3  # includes the shim's tower copies (~1 ms)
    res["reference_cpu_per_s"] = 1e3 / res["reference_cpu_ms"]
    res["cpu_cores"] = len(os.sched_getaffinity(0))
    b.close()
print(json.dumps(res))
