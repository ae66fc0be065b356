"""Focused check of the tensor-core base conversion (csrc/baseconv_tc.cu) on a GPU box.

  python tools/bconv_tc_check.py parity     -- ofhe_approx_switch_crt_basis against exact big-integer arithmetic
                                               (y_i = x_i * (Q/q_i)^-1 mod q_i; out_j = sum_i y_i * (Q/q_i mod p_j) mod p_j,
                                               dcrtpoly-impl.h:888-932) over shapes that exercise 1..5 accumulator chunks,
                                               split jobs (> 80 targets), exactly 16 source limbs, N = 128
  python tools/bconv_tc_check.py time       -- 15 -> 45 conversion at N = 2^16, CUDA events; run with OFHE_BCONV_TC=0 for
                                               the CUDA-core kernel
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
import torch  # noqa: E402

import ofhe_b200 as m  # noqa: E402


def columns(moduli, src, dst, x):
    """the 16 exact byte-digit column sums D_c[j][r] the tensor core must produce"""
    q = [int(moduli[i]) for i in src]
    Q = 1
    for v in q:
        Q *= v
    xs = x.astype(object)
    ys = [((xs[i] * pow((Q // qi) % qi, -1, qi)) % qi).astype(np.uint64) for i, qi in enumerate(q)]
    cols = np.zeros((16, len(dst), x.shape[1]), dtype=np.int64)
    for jj, j in enumerate(dst):
        p = int(moduli[j])
        for i, qi in enumerate(q):
            mm = (((Q // qi) % p) << 64) % p  # the matrix is held in Montgomery form (csrc/baseconv_tc.cu)
            for a in range(8):
                ya = ((ys[i] >> np.uint64(8 * a)) & np.uint64(255)).astype(np.int64)
                for b in range(8):
                    cols[a + b, jj] += ya * ((mm >> (8 * b)) & 255)
    return cols


def diagnose(ctx, ring, mod, src, dst, x, d, o):
    if os.environ.get("OFHE_BCONV_TC_DEBUG") is None:
        print("  (run with OFHE_BCONV_TC_DEBUG=1 for the per-column diagnosis)")
        return
    cols = columns(mod, src, dst, x)
    for c in range(16):
        os.environ["OFHE_BCONV_TC_COL"] = str(c + 1)
        ring.approx_switch_crt_basis(src, dst, d, o)
        torch.cuda.synchronize()
        got = ctx.download(o).astype(np.int64)
        w = np.argwhere(got != cols[c])
        msg = f"  column {c}: {len(w)} mismatches of {got.size}"
        if len(w):
            jj, rr = w[0]
            msg += f"; first at (limb {jj}, coeff {rr}): got {got[jj, rr]} want {cols[c][jj, rr]}"
            msg += f"; limbs hit {sorted(set(w[:, 0].tolist()))[:20]}; coeff%128 hit {sorted(set((w[:, 1] % 128).tolist()))[:12]}"
            msg += f"; row-multiset equal: {bool((np.sort(got[jj]) == np.sort(cols[c][jj])).all())}"
        print(msg, flush=True)
    os.environ["OFHE_BCONV_TC_COL"] = "0"


def expected(moduli, src, dst, x):
    q = [int(moduli[i]) for i in src]
    Q = 1
    for v in q:
        Q *= v
    xs = x.astype(object)
    out = []
    ys = []
    for i, qi in enumerate(q):
        hat = Q // qi
        ys.append((xs[i] * pow(hat % qi, -1, qi)) % qi)
    for j in dst:
        p = int(moduli[j])
        acc = 0
        for i, qi in enumerate(q):
            acc = acc + ys[i] * ((Q // qi) % p)
        out.append(acc % p)
    return np.array(out, dtype=object).astype(np.uint64)


def parity():
    ctx = m.Context(0)
    rng = np.random.default_rng(7)
    bad = 0
    for N, shapes in ((128, [(15, 45), (1, 3)]), (256, [(16, 17), (12, 90), (3, 5)]), (4096, [(15, 45), (8, 33)])):
        p = m.ckks_paramgen(N, 89, 59, 60, 3, 0)
        mod = p["moduli"]
        ring = m.Ring(ctx, N, mod, p["roots"])
        for ns, nd in shapes:
            src = list(range(2, 2 + ns))
            dst = [i for i in range(len(mod)) if i not in src][:nd]
            x = np.stack([rng.integers(0, int(mod[i]), N, dtype=np.uint64) for i in src])
            x[:, 0] = 0
            x[:, 1] = mod[src] - np.uint64(1)
            want = expected(mod, src, dst, x)
            d = ctx.upload(x)
            o = ctx.zeros(nd, N)
            ring.approx_switch_crt_basis(src, dst, d, o)
            torch.cuda.synchronize()
            got = ctx.download(o)
            ok = bool((got == want).all())
            bad += not ok
            print(f"N={N} {ns}->{nd}: {'ok' if ok else 'MISMATCH'}", flush=True)
            if not ok:
                w = np.argwhere(got != want)
                print("  first mismatches (limb, coeff):", w[:6].tolist(), "of", len(w))
                jj, rr = w[0]
                print(f"  got {int(got[jj, rr]):#x} want {int(want[jj, rr]):#x}")
                diagnose(ctx, ring, mod, src, dst, x, d, o)
        ring.close()
    ctx.close()
    print("PARITY", "FAILED" if bad else "PASSED")
    return bad


def timing():
    ctx = m.Context(0)
    N = 1 << 16
    p = m.ckks_paramgen(N, 44, 59, 60, 3, 0)
    ring = m.Ring(ctx, N, p["moduli"], p["roots"])
    src = list(range(15))
    dst = list(range(15, 60))
    d = ctx.zeros(15, N)
    ring.fill_uniform(d, src, 3)
    o = ctx.zeros(45, N)
    for _ in range(5):
        ring.approx_switch_crt_basis(src, dst, d, o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        ring.approx_switch_crt_basis(src, dst, d, o)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    alg = 8 * N * (15 + 45)
    print(json.dumps({"bconv_15_to_45_n16_us": us, "tc": os.environ.get("OFHE_BCONV_TC", "1"),
                      "algorithmic_gbs": alg / us / 1e3, "checksum": int(ctx.download(o).sum(dtype=np.uint64))}))
    ring.close()
    ctx.close()


if __name__ == "__main__":
    sys.exit(parity() if sys.argv[1] == "parity" else timing())
