"""BASELINE config 5: batched DCRTPoly::SwitchFormat sweep, N in 2^12..2^17 x L in {1,2,4,...,64}: achieved algorithmic
GB/s (16*N*L bytes per transform) vs the measured HBM peak, next to the reference CPU (oracle/_ref) on the same box.
  python tools/ntt_sweep.py [--cpu]      -> one JSON line per (N, L)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ofhe_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cpu", action="store_true", help="also time the reference CPU (needs oracle/_ref)")
ap.add_argument("--logn", type=int, nargs="*", default=[12, 13, 14, 15, 16, 17])
ap.add_argument("--limbs", type=int, nargs="*", default=[1, 2, 4, 8, 16, 32, 64])
a = ap.parse_args()
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6650.0
ctx = m.Context(0)
ref = None
if a.cpu:
    from oracle_lib import Reference, reference_available

    if reference_available():
        ref = Reference()
        ref.L.ref_set_threads(len(os.sched_getaffinity(0)))
for logn in a.logn:
    N = 1 << logn
    Lmax = max(a.limbs)
    q, roots = m.dcrt_params(N, Lmax, 60)  # ILDCRTParams(2N, L, 60) as poly-benchmark.h:86 builds them
    ring = m.Ring(ctx, N, q, roots)
    for L in a.limbs:
        # enough independent polynomials that the slab exceeds L2 (or at least 8 of them)
        batch = max(8, int(np.ceil(160e6 / (8.0 * N * L))))
        batch = min(batch, 256 // L) if L * batch > 256 else batch
        batch = max(batch, 1)
        rows = L * batch
        slab = ctx.zeros(rows, N)
        idx = list(range(L)) * batch
        ring.fill_uniform(slab, idx, 7)
        for _ in range(3):
            ring.ntt_fwd(slab, slab, idx)
            ring.ntt_inv(slab, slab, idx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            ring.ntt_fwd(slab, slab, idx)
            ring.ntt_inv(slab, slab, idx)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (2 * reps)  # per SwitchFormat of the whole slab
        gbs = 16.0 * N * rows / (us * 1e-6) / 1e9
        line = dict(N=N, L=L, batch=batch, us_per_dcrt_switchformat=us / batch, gbs=gbs, frac_of_measured_hbm=gbs / peak)
        if ref is not None:
            from oracle_lib import seeded_residues

            data = seeded_residues(3, q[:L], N)
            sec = ref.L.ref_time_switch_format(N, L, np.ascontiguousarray(q[:L]), np.ascontiguousarray(roots[:L]), data,
                                               max(2, int(2e5 / (N * L)) * 2))
            line["cpu_us_per_dcrt_switchformat"] = sec * 1e6
            line["cpu_threads"] = ref.L.ref_get_threads()
            line["speedup_vs_cpu"] = sec * 1e6 / (us / batch)
        print(json.dumps(line), flush=True)
    ring.close()
