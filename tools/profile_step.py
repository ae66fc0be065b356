"""One EvalMult step at BASELINE config 2 bracketed by cudaProfilerStart/Stop, for ncu (--profile-from-start off).
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \
      python tools/profile_step.py [--op evalmult|ntt|ntt240]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
import ofhe_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--op", default="evalmult")
ap.add_argument("--logn", type=int, default=16)
ap.add_argument("--depth", type=int, default=44)
a = ap.parse_args()
p = m.ckks_paramgen(1 << a.logn, a.depth, 59, 60, 0, 0)
N, sizeQ, sizeP, dnum = p["N"], p["sizeQ"], p["sizeP"], p["dnum"]
ctx = m.Context(0)
ring = m.Ring(ctx, N, p["moduli"], p["roots"])
hks = m.Hks(ring, sizeQ, sizeP, dnum)
evk = m.EvalKey(hks, seed=1)
idx2 = list(range(sizeQ)) * 2
ct1, ct2, out = ctx.zeros(2, sizeQ, N), ctx.zeros(2, sizeQ, N), ctx.zeros(2, sizeQ, N)
ring.fill_uniform(ct1.view(2 * sizeQ, N), idx2, 11)
ring.fill_uniform(ct2.view(2 * sizeQ, N), idx2, 12)
slab = ctx.zeros(sizeQ + sizeP, N)
lidx = list(range(sizeQ + sizeP))
ring.fill_uniform(slab, lidx, 5)
if a.op == "ntt240":  # exactly bench.py's roofline launch: forward NTT of a 240-limb slab (> L2)
    slab = ctx.zeros(240, N)
    lidx = [i % (sizeQ + sizeP) for i in range(240)]
    ring.fill_uniform(slab, lidx, 99)


def step():
    if a.op == "evalmult":
        hks.evalmult(sizeQ, ct1, ct2, evk, out)
    elif a.op == "ntt240":
        ring.ntt_fwd(slab, slab, lidx)
    else:
        ring.ntt_fwd(slab, slab, lidx)
        ring.ntt_inv(slab, slab, lidx)


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one", a.op)
