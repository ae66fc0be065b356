"""Side measurements printed in bench.py's "aux" object (single GPU): BASELINE.json configs 1, 3 and 5.

  config 1  DCRTPoly SwitchFormat fwd+inv, N = 4096, ONE 60-bit limb: latency of one transform pair (not a batch),
            eager launches and CUDA-graph replay, next to the reference's own SwitchFormat on one host core
            (what benchmark/src/poly-benchmark.h:283-299 `Native_avg_ntt_intt` times) -- bit-exact check included
  config 3  BFV EvalMult (HPSPOVERQ), N = 2^15, 12 Q + 12 R limbs, relinearised: mults/s and the algorithmic-byte
            roofline of the whole operation
  config 5  batched NTT sweep N in 2^12..2^17 x L in {1, 8, 64}: algorithmic GB/s (16*N*L per transform) vs HBM peak
All inputs synthetic; device timing with CUDA events on the launch stream after warm-up.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def _timeit(torch, f, reps, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps  # us


def config1(m, torch, ctx):
    N = 4096
    q, roots = m.dcrt_params(N, 1, 60)  # ILDCRTParams(2N, 1, 60) as poly-benchmark.h:86 builds it
    ring = m.Ring(ctx, N, q, roots)
    x = ctx.zeros(1, N)
    ring.fill_uniform(x, [0], 5)
    x0 = x.clone()

    def pair():
        ring.ntt_fwd(x, x, [0])
        ring.ntt_inv(x, x, [0])

    eager = _timeit(torch, pair, 200)
    assert bool((x == x0).all()), "SwitchFormat round trip is not the identity"
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        pair()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(10):
                pair()
        graph = _timeit(torch, g.replay, 50) / 10
    out = {"workload": "DCRTPoly SwitchFormat fwd+inv, N=4096, 1 x 60-bit limb (one polynomial, not a batch)",
           "gpu_us_per_fwd_inv_pair_eager": eager, "gpu_us_per_fwd_inv_pair_graph": graph}
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_lib import Reference, reference_available

        if reference_available():
            ref = Reference()
            ref.L.ref_set_threads(1)
            h = ctx.download(x0)[0].copy()
            want = h.copy()
            ref.L.ref_ntt_fwd(int(q[0]), int(roots[0]), N, want)
            y = x0.clone()
            ring.ntt_fwd(y, y, [0])
            out["bit_exact_vs_reference_forward"] = bool((ctx.download(y)[0] == want).all())
            sec = ref.L.ref_time_switch_format(N, 1, np.ascontiguousarray(q, dtype=np.uint64),
                                               np.ascontiguousarray(roots, dtype=np.uint64), h.reshape(1, N).copy(), 2000)
            out["reference_cpu_us_per_fwd_inv_pair_1_thread"] = 2 * sec * 1e6 if sec > 0 else None
            out["reference_note"] = "DCRTPoly::SwitchFormat of the unmodified reference, 1 thread (the operation " \
                                    "poly-benchmark's Native_avg_ntt_intt times)"
    except Exception as ex:
        out["reference_error"] = repr(ex)
    ring.close()
    return out


def config3(m, torch, ctx):
    g = np.load(os.path.join(ROOT, "tests", "golden", "bfv_cfg3_tables.npz"))
    N, nQ, nR = (int(v) for v in g["shape"])
    sizeP, dnum, _ = (int(v) for v in g["ks_shape"])
    T = {k[2:]: g[k] for k in g.files if k.startswith("T_")}
    ring = m.Ring(ctx, N, g["moduli"], g["roots"])
    mult = m.BfvMult(ring, nQ, nR, T)
    ring_ks = m.Ring(ctx, N, np.concatenate([g["moduli"][:nQ], g["moduliP"]]),
                     np.concatenate([g["roots"][:nQ], g["rootsP"]]))
    hks = m.Hks(ring_ks, nQ, sizeP, dnum)
    evk = m.EvalKey(hks, seed=3)
    ct1, ct2, out = ctx.zeros(2, nQ, N), ctx.zeros(2, nQ, N), ctx.zeros(2, nQ, N)
    ring.fill_uniform(ct1.view(2 * nQ, N), list(range(nQ)) * 2, 1)
    ring.fill_uniform(ct2.view(2 * nQ, N), list(range(nQ)) * 2, 2)
    us = _timeit(torch, lambda: mult.evalmult(hks, ct1, ct2, evk, out), 30)
    limb = 8.0 * N
    nQR = nQ + nR
    # algorithmic bytes (SURVEY 8 d accounting, unfused): 4 polys INTT (Q) + expand Q->R (read Q write R) + 4 polys NTT
    # over QR + tensor (4 in + 3 out over QR) + 3 polys INTT over QR + scale-and-round QR->Q + 3 polys NTT over Q
    # + hybrid key switch of c2 over Q u P + final adds
    nb = limb * (4 * 2 * nQ + 4 * (nQ + nR) + 4 * 2 * nQR + 7 * nQR + 3 * 2 * nQR + 3 * (nQR + nQ) + 3 * 2 * nQ)
    alpha = (nQ + dnum - 1) // dnum
    nb += limb * (2 * nQ + dnum * ((alpha + (nQ + sizeP - alpha)) + 2 * (nQ + sizeP - alpha))
                  + (dnum * (nQ + sizeP) + 2 * dnum * (nQ + sizeP) + 2 * (nQ + sizeP))
                  + 2 * (2 * sizeP + (sizeP + nQ) + 2 * nQ + 3 * nQ) + 2 * 3 * nQ)
    peak = _peak()
    res = {"workload": "BFV EvalMult HPSPOVERQ + relinearisation, N=2^15, 12 Q + 12 R limbs (60-bit), HYBRID 4 P limbs dnum 3",
           "evalmult_relin_us": us, "evalmult_relin_per_s": 1e6 / us,
           "roofline": {"bound": "hbm", "algorithmic_bytes": nb, "achieved_gbs": nb / us * 1e-3, "peak_gbs": peak,
                        "frac": nb / us * 1e-3 / peak}}
    for o in (evk, hks, mult, ring_ks, ring):
        o.close()
    return res


def config5(m, torch, ctx):
    peak = _peak()
    rows = []
    for logn in (12, 13, 14, 15, 16, 17):
        N = 1 << logn
        q, roots = m.dcrt_params(N, 64, 60)
        ring = m.Ring(ctx, N, q, roots)
        for L in (1, 8, 64):
            batch = max(1, min(256 // L, int(np.ceil(160e6 / (8.0 * N * L)))))
            n = L * batch
            slab = ctx.zeros(n, N)
            idx = sorted(list(range(L)) * batch)  # modulus-major
            ring.fill_uniform(slab, idx, 7)

            def pair():
                ring.ntt_fwd(slab, slab, idx)
                ring.ntt_inv(slab, slab, idx)

            us = _timeit(torch, pair, 5, warm=2)
            gbs = 2 * 16.0 * N * n / us * 1e-3
            rows.append({"logN": logn, "L": L, "batch": batch, "us_per_dcrt_switchformat": us / (2 * batch),
                         "gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4)})
            del slab
        ring.close()
    return {"workload": "batched DCRTPoly SwitchFormat (fwd + inv), 16*N*L algorithmic bytes per transform, slab > L2 "
                        "where 256 limbs allow", "peak_gbs": peak, "rows": rows}


def config2_resident(m, torch, ctx):
    """SURVEY 8 f1/f2 at the SIZE of BASELINE config 2 (N = 2^16, 45 + 15 limbs, dnum 3): device-resident chains and
    hoisted rotations, eager and as one CUDA graph.  Synthetic residues and keys (throughput does not depend on them);
    the same entry points are parity-tested bit-exact at N = 4096 (tests/test_gpu_chain.py, test_gpu_parity.py)."""
    p = m.ckks_paramgen(1 << 16, 44, 59, 60, 3, 0)
    N, L, sizeP = p["N"], p["sizeQ"], p["sizeP"]
    ring = m.Ring(ctx, N, p["moduli"], p["roots"])
    hks = m.Hks(ring, L, sizeP, p["dnum"])
    evk, rk = m.EvalKey(hks, seed=11), m.EvalKey(hks, seed=12)
    ct1, ct2 = ctx.zeros(2, L, N), ctx.zeros(2, L, N)
    ring.fill_uniform(ct1.view(2 * L, N), list(range(L)) * 2, 21)
    ring.fill_uniform(ct2.view(2 * L, N), list(range(L)) * 2, 22)
    b_m, b_r = ctx.zeros(2, L, N), ctx.zeros(2, L - 1, N)
    b_rot, b_s = ctx.zeros(2, L - 1, N), ctx.zeros(2, L - 1, N)
    k = m.find_automorphism_index(1, 2 * N)
    idx2 = list(range(L - 1)) * 2
    for lvl in (L, L - 1):
        hks.reserve(lvl)

    def chain():  # EvalMult -> Rescale -> EvalRotate -> EvalAdd, nothing leaves the device
        hks.evalmult(L, ct1, ct2, evk, b_m)
        hks.rescale(L, b_m, b_r, 2)
        hks.rotate(L - 1, k, b_r, rk, b_rot)
        ring.binop(0, b_s.view(2 * (L - 1), N), b_rot.view(2 * (L - 1), N), b_r.view(2 * (L - 1), N), idx2)

    out = {"workload": "CKKS chain EvalMult+relin -> Rescale -> EvalRotate (own key) -> EvalAdd, N=2^16, 45+15 limbs, "
                       "device-resident (SURVEY 8 f1)"}
    out["chain_us_eager"] = _timeit(torch, chain, 10)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        chain()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            chain()
        out["chain_us_graph"] = _timeit(torch, g.replay, 10)
    out["rescale_us"] = _timeit(torch, lambda: hks.rescale(L, b_m, b_r, 2), 20)
    out["rotate_us"] = _timeit(torch, lambda: hks.rotate(L - 1, k, b_r, rk, b_rot), 20)
    out["rotations_per_s"] = 1e6 / out["rotate_us"]
    # hoisted rotations: one ModUp shared by R rotations of the same ciphertext
    R = 8
    nd = hks.num_digits(L - 1)
    digits = ctx.zeros(nd, L - 1 + sizeP, N)
    outs = [ctx.zeros(2, L - 1, N) for _ in range(2)]

    def hoisted():
        hks.modup(L - 1, b_r[1], digits)
        for j in range(R):
            hks.fast_rotation(L - 1, k, b_r, digits, rk, outs[j & 1])

    us = _timeit(torch, hoisted, 5)
    out["hoisted_8_rotations_us"] = us
    out["hoisted_rotations_per_s"] = R * 1e6 / us
    # one level of the double-hoisted BSGS linear transform (ofhe_ckks_bsgs_level, ckksrns-fhe.cpp:1934-1987):
    # g = b = 8 (64 diagonals in QlP, 8 baby-step + 7 giant-step hoisted rotations, 9 ModDowns), synthetic keys
    g = b = 8
    Ls = L - 1
    keys_in = [None] + [m.EvalKey(hks, seed=30 + j) for j in range(1, g)]
    keys_out = [None] + [m.EvalKey(hks, seed=50 + i) for i in range(1, b)]
    k_in = [0] + [m.find_automorphism_index(j, 2 * N) for j in range(1, g)]
    k_out = [0] + [m.find_automorphism_index(g * i, 2 * N) for i in range(1, b)]
    A = ctx.zeros(b * g, Ls + sizeP, N)
    for d in range(b * g):  # one diagonal (Ls + sizeP limbs) per call: a launch addresses at most 256 limbs
        ring.fill_uniform(A[d], list(range(Ls)) + list(range(L, L + sizeP)), 77 + d)
    ws = hks.bsgs_workspace(Ls, g)
    res = ctx.zeros(2, Ls, N)
    us = _timeit(torch, lambda: hks.bsgs_level(Ls, k_in, k_out, keys_in, keys_out, A, b * g, b_r, res, ws), 3, warm=1)
    out["bsgs_level_g8_b8_us"] = us
    out["bsgs_level_note"] = "64 plaintext diagonals, 15 hoisted rotations, device-resident accumulation in QlP (SURVEY 8 f2)"
    for o in [k for k in keys_in + keys_out if k is not None] + [evk, rk, hks, ring]:
        o.close()
    return out


def config5_sharded(m, torch, ctx, dist, rank, world):
    """BASELINE config 5 at N GPUs: the 256 limbs of every (N, L = 64, batch = 4) point are dealt round-robin to the
    ranks (RNS limbs are independent: no exchange), every rank transforms its share, the time of a point is the MAX over
    ranks (CUDA events, all_reduce), the throughput is the whole point's 16*N*256 bytes per direction over that time."""
    peak = _peak()
    rows = []
    for logn in (12, 14, 16, 17):
        N = 1 << logn
        q, roots = m.dcrt_params(N, 64, 60)
        ring = m.Ring(ctx, N, q, roots)
        total = 256 if logn < 17 else 192
        mine = [i % 64 for i in range(total) if i % world == rank]
        idx = sorted(mine)
        slab = ctx.zeros(len(idx), N)
        ring.fill_uniform(slab, idx, 7)

        def pair():
            ring.ntt_fwd(slab, slab, idx)
            ring.ntt_inv(slab, slab, idx)

        dist.barrier()
        us = torch.tensor([_timeit(torch, pair, 5, warm=2)], device=f"cuda:{ctx.device}", dtype=torch.float64)
        dist.all_reduce(us, op=dist.ReduceOp.MAX)
        gbs = 2 * 16.0 * N * total / float(us.item()) * 1e-3
        rows.append({"logN": logn, "L": 64, "limbs_total": total, "limbs_per_rank": len(idx), "us_fwd_inv_pair_max_rank": float(us.item()),
                     "gbs_all_ranks": round(gbs, 1), "frac_of_measured_hbm_x_ranks": round(gbs / (peak * world), 4)})
        del slab
        ring.close()
    return {"workload": "batched DCRTPoly SwitchFormat (fwd + inv) with the limbs dealt to the ranks (no exchange), "
                        "16*N*L algorithmic bytes per transform", "n_gpus": world, "peak_gbs_per_gpu": peak, "rows": rows}


def base_conversion(m, torch, ctx):
    """RNS fast base conversion 15 -> 45 limbs at N = 2^16 (one ModUp digit of config 2), the tensor-core kernel
    (csrc/baseconv_tc.cu) unless OFHE_BCONV_TC=0: algorithmic bytes 8*N*(15+45)."""
    N = 1 << 16
    p = m.ckks_paramgen(N, 44, 59, 60, 3, 0)
    ring = m.Ring(ctx, N, p["moduli"], p["roots"])
    src, dst = list(range(15)), list(range(15, 60))
    d, o = ctx.zeros(15, N), ctx.zeros(45, N)
    ring.fill_uniform(d, src, 3)
    us = _timeit(torch, lambda: ring.approx_switch_crt_basis(src, dst, d, o), 30, warm=5)
    ring.close()
    alg = 8.0 * N * 60
    return {"workload": "ApproxSwitchCRTBasis 15 -> 45 limbs, N=2^16 (QHatInv multiply included)", "us": us,
            "kernel": "bconv_tc_kernel (tcgen05.mma kind::i8, TMEM accumulators)" if os.environ.get("OFHE_BCONV_TC", "1") != "0"
            else "bconv_kernel (CUDA cores)",
            "algorithmic_gbs": alg / us / 1e3, "frac_of_measured_hbm": alg / us / 1e3 / _peak(),
            "cuda_core_kernel_us_committed": 50.2, "source": "profiles/r02s_kernels_tc_vs_cuda.json"}


def run(m, torch, ctx):
    out = {}
    for name, fn in (("config1_ntt4096_single_limb", config1), ("config2_resident_chain_and_rotations", config2_resident),
                     ("config3_bfv_evalmult_n15", config3), ("base_conversion_15_45", base_conversion),
                     ("config5_ntt_sweep", config5)):
        try:
            out[name] = fn(m, torch, ctx)
        except Exception as ex:
            out[name] = {"error": repr(ex)}
    return out
