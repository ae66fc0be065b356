"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference (oracle/_ref, built by
oracle/Makefile.ref from /root/reference).  Run in the build container:  python tests/golden/make_golden.py
The vectors pin the oracle (and through it the CUDA path) on boxes where the reference build is absent.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import Reference, seeded_residues  # noqa: E402


def main():
    r = Reference()
    L = r.L
    # ---- NTT: N = 64 and N = 4096 (one 60-bit limb each), N = 8 with two 28-bit limbs (UnitTestNTT.cpp:92-119 shape)
    out = {}
    for tag, N, bits, nl in (("n64", 64, 60, 2), ("n4096", 4096, 60, 1), ("n8", 8, 28, 2)):
        q = [L.ref_last_prime(bits, 2 * N)]
        for _ in range(nl - 1):
            q.append(L.ref_previous_prime(q[-1], 2 * N))
        q = np.array(q, dtype=np.uint64)
        roots = np.array([L.ref_root_of_unity(2 * N, int(x)) for x in q], dtype=np.uint64)
        x = seeded_residues(20260922, q, N)
        y = r.dcrt_switch_format(q, roots, x, from_eval=False)
        out[f"{tag}_moduli"], out[f"{tag}_roots"], out[f"{tag}_coeff"], out[f"{tag}_eval"] = q, roots, x, y
    T, Tp, Ti, Tip, ni, nip = r.twiddles(int(out["n64_moduli"][0]), int(out["n64_roots"][0]), 64)
    out.update(n64_T=T, n64_Tp=Tp, n64_Ti=Ti, n64_Tip=Tip, n64_nInv=np.array([ni, nip], dtype=np.uint64))
    # limb-wise ops + automorphism + switch modulus on the n64 ring
    q, roots = out["n64_moduli"], out["n64_roots"]
    a, b = seeded_residues(1, q, 64), seeded_residues(2, q, 64)
    for op, name in ((0, "add"), (1, "sub"), (2, "mul"), (3, "neg")):
        out[f"n64_{name}"] = r.dcrt_binop(q, roots, op, a, b)
    out["n64_a"], out["n64_b"] = a, b
    out["n64_auto5"] = r.dcrt_automorphism(q, roots, 5, a)
    out["n64_auto127"] = r.dcrt_automorphism(q, roots, 127, a)
    out["n64_swmod_down"] = r.switch_modulus(int(q[0]), int(q[1]), a[0])
    out["n64_swmod_up"] = r.switch_modulus(int(q[1]), int(q[0]), b[1])
    np.savez_compressed(os.path.join(HERE, "core_vectors.npz"), **out)

    # ---- CKKS hybrid key switching: N = 64, depth 4, 40/50-bit moduli, dnum 2 (uneven last digit at lower levels)
    for tag, args in (("ckks_n64", (64, 4, 40, 50, 2, 0)), ("ckks_n128_flex", (128, 5, 45, 55, 3, 2))):
        c = r.ckks(*args)
        g = dict(args=np.array(args, dtype=np.int64), shape=np.array([c.N, c.sizeQ, c.sizeP, c.dnum, c.alpha]),
                 moduli=c.moduli, roots=c.roots, PInvModq=c.PInvModq(), PModq=c.PModq(),
                 PHatInvModp=c.PHatInvModp(), PHatModq=c.PHatModq())
        ea, eb = c.evalmult_key()
        g["evk_a"], g["evk_b"] = ea, eb
        for sizeQl in (c.sizeQ, c.sizeQ - 1, 1):
            p = seeded_residues(100 + sizeQl, c.moduli[:sizeQl], c.N)
            g[f"l{sizeQl}_c"] = p
            g[f"l{sizeQl}_digits"] = c.modup(sizeQl, p)
            k0, k1 = c.keyswitch_core(sizeQl, p)
            g[f"l{sizeQl}_ks"] = np.stack([k0, k1])
            ct1 = np.stack([seeded_residues(200 + sizeQl, c.moduli[:sizeQl], c.N),
                            seeded_residues(300 + sizeQl, c.moduli[:sizeQl], c.N)])
            ct2 = np.stack([seeded_residues(400 + sizeQl, c.moduli[:sizeQl], c.N),
                            seeded_residues(500 + sizeQl, c.moduli[:sizeQl], c.N)])
            g[f"l{sizeQl}_ct1"], g[f"l{sizeQl}_ct2"] = ct1, ct2
            g[f"l{sizeQl}_evalmult"] = c.evalmult(sizeQl, ct1, ct2)
            if sizeQl > 1:
                g[f"l{sizeQl}_rescale"] = c.rescale(sizeQl, ct1)
                A, B = c.rescale_tables(sizeQl)
                g[f"l{sizeQl}_rescaleA"], g[f"l{sizeQl}_rescaleB"] = A, B
            for part in range(c.num_digits(sizeQl)):
                ds = c.digit_size(sizeQl, part)
                g[f"l{sizeQl}_p{part}_hatinv"] = c.PartQlHatInvModq(part, ds)
                m, cm = c.PartQlHatModp(sizeQl, part)
                g[f"l{sizeQl}_p{part}_hatmodp"], g[f"l{sizeQl}_p{part}_compl"] = m, cm
        k, ra, rb = c.rotate_key(1)
        g["rot_k"], g["rot_a"], g["rot_b"] = np.array([k]), ra, rb
        sizeQl = c.sizeQ - 1
        g["rot_out"] = c.rotate(sizeQl, 1, g[f"l{sizeQl}_ct1"])
        # hoisted rotations: EvalFastRotation / EvalFastRotationExt / KeySwitchExt / KeySwitchDown
        g["rot_fast"] = c.fast_rotate(sizeQl, 1, g[f"l{sizeQl}_ct1"])
        g["rot_ext1"] = c.fast_rotate_ext(sizeQl, 1, True, g[f"l{sizeQl}_ct1"])
        g["rot_ext0"] = c.fast_rotate_ext(sizeQl, 1, False, g[f"l{sizeQl}_ct1"])
        g["ks_ext1"] = c.keyswitch_ext(sizeQl, True, g[f"l{sizeQl}_ct1"])
        g["ks_ext0"] = c.keyswitch_ext(sizeQl, False, g[f"l{sizeQl}_ct1"])
        g["ks_down"] = c.keyswitch_down(sizeQl, g["rot_ext1"])
        # an honest encryption / multiplication / decryption, for the end-to-end sanity test
        vals1 = np.array([0.5, -0.25, 1.0, 0.125])
        vals2 = np.array([2.0, 4.0, -1.0, 8.0])
        e1, e2 = c.encrypt(vals1), c.encrypt(vals2)
        g["enc1"], g["enc2"] = e1, e2
        prod = c.evalmult(c.sizeQ, e1, e2)
        g["enc_prod"] = prod
        g["dec_prod"] = c.decrypt(c.sizeQ, 2, prod, 4)
        np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **g)
        c.close()
    # ---- BFV HPSPOVERQ (rows a12 / a13, config 3): a small context with every input / output, and the TABLES ONLY of
    #      the config-3 shape (N = 2^15, 60-bit limbs, log2 q ~ 720 -> 12 Q limbs) for full-size oracle-vs-GPU tests
    for tag, args, with_data in (("bfv_n64", (64, 2, 60, 65537, 0), True), ("bfv_n128_d4", (128, 4, 50, 786433, 0), True),
                                 ("bfv_cfg3_tables", (1 << 15, 20, 60, 65537, 0), False)):
        b = r.bfv(*args)
        g = dict(args=np.array(args, dtype=np.int64), shape=np.array([b.N, b.sizeQ, b.sizeR]), moduli=b.moduli,
                 roots=b.roots)
        if not with_data:
            ks = b.ks_shape()
            g.update(ks_shape=np.array([ks["sizeP"], ks["dnum"], ks["alpha"]]), moduliP=ks["moduliP"], rootsP=ks["rootsP"])
        for k, v in b.tables().items():
            g["T_" + k] = v
        if with_data:
            Q = b.moduli[:b.sizeQ]
            xe, xc = seeded_residues(600, Q, b.N), seeded_residues(601, Q, b.N)
            xq = seeded_residues(602, b.moduli, b.N)
            # edge coefficients: zeros, q-1 everywhere, and a value making the overflow estimate land near an integer
            xc[:, 0] = 0
            xc[:, 1] = Q - np.uint64(1)
            xq[:, 0] = 0
            xq[:, 1] = b.moduli - np.uint64(1)
            g.update(expand_in=xe, expand_out=b.expand_crt_basis(xe), switch_in=xc, switch_out=b.switch_crt_basis(xc),
                     fastexp_in=xc, fastexp_out=b.fast_expand_p_over_q(xc), sr_in=xq, sr_out=b.scale_and_round(xq))
            ct1 = np.stack([seeded_residues(603, Q, b.N), seeded_residues(604, Q, b.N)])
            ct2 = np.stack([seeded_residues(605, Q, b.N), seeded_residues(606, Q, b.N)])
            g.update(ct1=ct1, ct2=ct2, evalmult_norelin=b.evalmult_norelin(ct1, ct2))
            # relinearised product (HYBRID key switch of the BFV context): P moduli, key, cc->EvalMult output
            ks = b.ks_shape()
            ea, eb = b.evalmult_key()
            g.update(ks_shape=np.array([ks["sizeP"], ks["dnum"], ks["alpha"]]), moduliP=ks["moduliP"], rootsP=ks["rootsP"],
                     evk_a=ea, evk_b=eb, evalmult=b.evalmult(ct1, ct2))
        print(tag, "sizeQ", b.sizeQ, "sizeR", b.sizeR)
        np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **g)
        b.close()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
