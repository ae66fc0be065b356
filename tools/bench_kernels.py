"""Kernel-level timings at config-2 shapes (CUDA events on the launch stream, inputs larger than L2 where relevant).
  python tools/bench_kernels.py            -> one JSON dict per line
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
import ofhe_b200 as m  # noqa: E402

logn = int(os.environ.get("LOGN", "16"))
p = m.ckks_paramgen(1 << logn, 44, 59, 60, 0, 0)
N, sizeQ, sizeP, dnum = p["N"], p["sizeQ"], p["sizeP"], p["dnum"]
ctx = m.Context(0)
ring = m.Ring(ctx, N, p["moduli"], p["roots"])
hks = m.Hks(ring, sizeQ, sizeP, dnum)
evk = m.EvalKey(hks, seed=1)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


L = 240
slab = ctx.zeros(L, N)
lidx = [i % (sizeQ + sizeP) for i in range(L)]
ring.fill_uniform(slab, lidx, 3)
res = {}
res["ntt_fwd_us_per_limb"] = timeit(lambda: ring.ntt_fwd(slab, slab, lidx)) / L
res["ntt_inv_us_per_limb"] = timeit(lambda: ring.ntt_inv(slab, slab, lidx)) / L
src = ctx.zeros(15, N)
ring.fill_uniform(src, list(range(15)), 4)
dst = ctx.zeros(45, N)
dst_idx = list(range(15, 60))
res["bconv_15to45_us"] = timeit(lambda: ring.approx_switch_crt_basis(list(range(15)), dst_idx, src, dst))
a, b, o = ctx.zeros(45, N), ctx.zeros(45, N), ctx.zeros(45, N)
idx = list(range(45))
ring.fill_uniform(a, idx, 5)
ring.fill_uniform(b, idx, 6)
res["ew_mul_45limbs_us"] = timeit(lambda: ring.binop(2, o, a, b, idx))
res["ew_add_45limbs_us"] = timeit(lambda: ring.binop(0, o, a, b, idx))
ct1, ct2, out = ctx.zeros(2, sizeQ, N), ctx.zeros(2, sizeQ, N), ctx.zeros(2, sizeQ, N)
ring.fill_uniform(ct1.view(2 * sizeQ, N), idx * 2, 7)
ring.fill_uniform(ct2.view(2 * sizeQ, N), idx * 2, 8)
res["evalmult_us"] = timeit(lambda: hks.evalmult(sizeQ, ct1, ct2, evk, out), reps=10)
digits = ctx.zeros(3, 60, N)
res["modup_us"] = timeit(lambda: hks.modup(sizeQ, ct1[1], digits), reps=10)
e01 = ctx.zeros(2, 60, N)
res["inner_product_us"] = timeit(lambda: hks.inner_product(sizeQ, digits, evk, e01), reps=10)
md = ctx.zeros(2, sizeQ, N)
res["moddown2_us"] = timeit(lambda: hks.approx_mod_down(sizeQ, e01, md, 2), reps=10)
res["env"] = {k: v for k, v in os.environ.items() if k.startswith("OFHE_")}
print(json.dumps(res))
