"""BFV rows of SURVEY.md section 8: a12 (exact SwitchCRTBasis / ExpandCRTBasis / FastExpandCRTBasisPloverQ) and a13
(ScaleAndRound), the `double`-assisted conversions of HPSPOVERQ EvalMult (BASELINE config 3).
CPU tier: oracle vs golden vectors from the reference and vs the live reference.  GPU tier: CUDA vs oracle / golden,
plus the composed EvalMultNoRelin pipeline at the golden sizes and the a12/a13 kernels at the config-3 size."""
import numpy as np
import pytest

from oracle_lib import seeded_residues


def _T(g):
    return {k[2:]: g[k] for k in g.files if k.startswith("T_")}


def oracle_expand(o, g, T, xe):
    sizeQ = int(g["shape"][1])
    Q, R, roots = g["moduli"][:sizeQ], g["moduli"][sizeQ:], g["roots"]
    coef = o.dcrt_switch_format(Q, roots[:sizeQ], xe, True)
    pr = o.switch_crt_basis(Q, T["QlHatInvModq"], R, T["QlHatModr"], T["alphaQlModr"], T["qInv"], coef)
    return np.concatenate([xe, o.dcrt_switch_format(R, roots[sizeQ:], pr, False)])


def oracle_fast_expand(o, g, T, xc):
    sizeQ = int(g["shape"][1])
    Q, R = g["moduli"][:sizeQ], g["moduli"][sizeQ:]
    partR = o.approx_switch_crt_basis(Q, T["mNegRlQlHatInvModq"], R, T["qInvModr"], xc)
    partQ = o.switch_crt_basis(R, T["RlHatInvModr"], Q, T["RlHatModq"], T["alphaRlModq"], T["rInv"], partR)
    return np.concatenate([partQ, partR])


@pytest.mark.parametrize("name", ["bfv_n64", "bfv_n128_d4"])
def test_oracle_vs_golden(oracle, golden, name):
    g = golden(name)
    T = _T(g)
    sizeQ = int(g["shape"][1])
    Q, R = g["moduli"][:sizeQ], g["moduli"][sizeQ:]
    got = oracle.switch_crt_basis(Q, T["QlHatInvModq"], R, T["QlHatModr"], T["alphaQlModr"], T["qInv"], g["switch_in"])
    assert (got == g["switch_out"]).all()
    assert (oracle_expand(oracle, g, T, g["expand_in"]) == g["expand_out"]).all()
    assert (oracle_fast_expand(oracle, g, T, g["fastexp_in"]) == g["fastexp_out"]).all()
    sr = oracle.scale_and_round(Q, True, T["tQlSlHatInvModsDivsModq"], T["tQlSlHatInvModsDivsFrac"], g["sr_in"])
    assert (sr == g["sr_out"]).all()
    # the composed HPSPOVERQ EvalMult without relinearisation (bfvrns-leveledshe.cpp:198-439)
    ct1, ct2 = g["ct1"], g["ct2"]
    moduli, roots = g["moduli"], g["roots"]
    e1 = [oracle_expand(oracle, g, T, ct1[k]) for k in range(2)]
    e2 = []
    for k in range(2):
        coef = oracle.dcrt_switch_format(Q, roots[:sizeQ], ct2[k], True)
        e2.append(oracle.dcrt_switch_format(moduli, roots, oracle_fast_expand(oracle, g, T, coef), False))
    mul = lambda a, b: oracle.dcrt_binop(moduli, 2, a, b)
    add = lambda a, b: oracle.dcrt_binop(moduli, 0, a, b)
    cv = [mul(e1[0], e2[0]), add(mul(e1[0], e2[1]), mul(e1[1], e2[0])), mul(e1[1], e2[1])]
    for k in range(3):
        coef = oracle.dcrt_switch_format(moduli, roots, cv[k], True)
        out = oracle.scale_and_round(Q, True, T["tQlSlHatInvModsDivsModq"], T["tQlSlHatInvModsDivsFrac"], coef)
        assert (out == g["evalmult_norelin"][k]).all(), k


def oracle_evalmult_relin(oracle, g):
    """RelinearizeCore on the reference's own 3-element product (bfvrns-leveledshe.cpp:888-935): NTT, hybrid
    KeySwitchCore of c2 over BFV's Q and P, c0 += ab0, c1 += ab1.  Returns [2][sizeQ][N] EVALUATION."""
    N, sizeQ, _ = (int(v) for v in g["shape"])
    sizeP, dnum, alpha = (int(v) for v in g["ks_shape"])
    Q, rQ = g["moduli"][:sizeQ], g["roots"][:sizeQ]
    h = oracle.hks_from_moduli(N, sizeQ, sizeP, dnum, np.concatenate([Q, g["moduliP"]]), np.concatenate([rQ, g["rootsP"]]))
    try:
        assert h.alpha == alpha
        c = [oracle.dcrt_switch_format(Q, rQ, g["evalmult_norelin"][k], False) for k in range(3)]
        ab0, ab1 = h.keyswitch_core(sizeQ, c[2], g["evk_a"], g["evk_b"])
        return np.stack([oracle.dcrt_binop(Q, 0, c[0], ab0), oracle.dcrt_binop(Q, 0, c[1], ab1)])
    finally:
        h.close()


@pytest.mark.parametrize("name", ["bfv_n64", "bfv_n128_d4"])
def test_oracle_relinearised_product_golden(oracle, golden, name):
    g = golden(name)
    assert (oracle_evalmult_relin(oracle, g) == g["evalmult"]).all()


class _Tables(dict):
    """dict with the `.files` attribute of an npz archive, for contexts created live from the reference"""

    @property
    def files(self):
        return list(self.keys())


def _live_tables(b):
    g = _Tables(shape=np.array([b.N, b.sizeQ, b.sizeR]), moduli=b.moduli, roots=b.roots)
    for k, v in b.tables().items():
        g["T_" + k] = v
    return g


def test_oracle_vs_reference_live(oracle, reference):
    b = reference.bfv(1 << 9, 3, 55, 65537, 0)
    try:
        g = _live_tables(b)
        T = _T(g)
        Q, R = b.moduli[:b.sizeQ], b.moduli[b.sizeQ:]
        xc = seeded_residues(1, Q, b.N)
        assert (oracle.switch_crt_basis(Q, T["QlHatInvModq"], R, T["QlHatModr"], T["alphaQlModr"], T["qInv"], xc) ==
                b.switch_crt_basis(xc)).all()
        assert (oracle_expand(oracle, g, T, xc) == b.expand_crt_basis(xc)).all()
        assert (oracle_fast_expand(oracle, g, T, xc) == b.fast_expand_p_over_q(xc)).all()
        xq = seeded_residues(2, b.moduli, b.N)
        assert (oracle.scale_and_round(Q, True, T["tQlSlHatInvModsDivsModq"], T["tQlSlHatInvModsDivsFrac"], xq) ==
                b.scale_and_round(xq)).all()
    finally:
        b.close()


# ------------------------------------------------------------------------------------------------------ GPU tier
class DeviceBfv:
    """HPSPOVERQ EvalMult plumbing over the C ABI (scheme logic stays host side; every DCRTPoly call is a device call)."""

    def __init__(self, gpu, g):
        self.ctx, self.m = gpu
        m = self.m
        self.g, self.T = g, _T(g)
        self.N, self.sizeQ, self.sizeR = (int(v) for v in g["shape"])
        self.ring = m.Ring(self.ctx, self.N, g["moduli"], g["roots"])
        T, nQ, nR = self.T, self.sizeQ, self.sizeR
        self.iq, self.ir, self.iqr = list(range(nQ)), list(range(nQ, nQ + nR)), list(range(nQ + nR))
        B = m.BaseConv
        self.q_to_r = B.switch_crt_basis(self.ring, self.iq, T["QlHatInvModq"], self.ir, T["QlHatModr"], True,
                                         T["alphaQlModr"], T["qInv"])
        self.q_to_r_approx = B.switch_crt_basis(self.ring, self.iq, T["mNegRlQlHatInvModq"], self.ir, T["qInvModr"], False)
        self.r_to_q = B.switch_crt_basis(self.ring, self.ir, T["RlHatInvModr"], self.iq, T["RlHatModq"], True,
                                         T["alphaRlModq"], T["rInv"])
        self.sr = B.scale_and_round(self.ring, self.ir, self.iq, True, T["tQlSlHatInvModsDivsModq"],
                                    T["tQlSlHatInvModsDivsFrac"])

    def close(self):
        for o in (self.q_to_r, self.q_to_r_approx, self.r_to_q, self.sr, self.ring):
            o.close()

    def expand(self, x_eval_d):  # ExpandCRTBasis, EVALUATION in / out  (dcrtpoly-impl.h:1088-1116)
        ctx, nQ, nR, N = self.ctx, self.sizeQ, self.sizeR, self.N
        out = ctx.zeros(nQ + nR, N)
        coef = ctx.zeros(nQ, N)
        self.ring.ntt_inv(x_eval_d, coef, self.iq)
        self.q_to_r.apply(coef, out[nQ:])
        self.ring.ntt_fwd(out[nQ:], out[nQ:], self.ir)
        out[:nQ].copy_(x_eval_d)
        return out

    def fast_expand(self, x_coef_d):  # FastExpandCRTBasisPloverQ, COEFFICIENT in / out  (:1151-1164)
        out = self.ctx.zeros(self.sizeQ + self.sizeR, self.N)
        self.q_to_r_approx.apply(x_coef_d, out[self.sizeQ:])
        self.r_to_q.apply(out[self.sizeQ:], out[:self.sizeQ])
        return out

    def evalmult_norelin(self, ct1, ct2):
        ctx, nQ, N = self.ctx, self.sizeQ, self.N
        d1, d2 = ctx.upload(ct1), ctx.upload(ct2)
        e1 = [self.expand(d1[k]) for k in range(2)]
        e2 = []
        for k in range(2):
            coef = ctx.zeros(nQ, N)
            self.ring.ntt_inv(d2[k], coef, self.iq)
            f = self.fast_expand(coef)
            self.ring.ntt_fwd(f, f, self.iqr)
            e2.append(f)
        cv = [ctx.zeros(nQ + self.sizeR, N) for _ in range(3)]
        self.ring.binop(2, cv[0], e1[0], e2[0], self.iqr)
        self.ring.binop(2, cv[1], e1[0], e2[1], self.iqr)
        self.ring.binop(4, cv[1], e1[1], e2[0], self.iqr)
        self.ring.binop(2, cv[2], e1[1], e2[1], self.iqr)
        outs = []
        for k in range(3):
            self.ring.ntt_inv(cv[k], cv[k], self.iqr)
            o = ctx.zeros(nQ, N)
            self.sr.apply(cv[k], o)
            outs.append(ctx.download(o).copy())
        return np.stack(outs)


class ProductBfv:
    """The library's own BFV multiplication (csrc/bfv.cu) plus the HYBRID context over Q + P for relinearisation."""

    def __init__(self, gpu, g, ks=None):
        self.ctx, m = gpu
        self.N, self.sizeQ, self.sizeR = (int(v) for v in g["shape"])
        self.ring = m.Ring(self.ctx, self.N, g["moduli"], g["roots"])
        self.mult = m.BfvMult(self.ring, self.sizeQ, self.sizeR, _T(g))
        self.hks = self.ring_ks = None
        ks = ks or ({"sizeP": int(g["ks_shape"][0]), "dnum": int(g["ks_shape"][1]), "moduliP": g["moduliP"],
                     "rootsP": g["rootsP"]} if "ks_shape" in g.files else None)
        if ks:
            nQ = self.sizeQ
            self.ring_ks = m.Ring(self.ctx, self.N, np.concatenate([g["moduli"][:nQ], ks["moduliP"]]),
                                  np.concatenate([g["roots"][:nQ], ks["rootsP"]]))
            self.hks = m.Hks(self.ring_ks, nQ, ks["sizeP"], ks["dnum"])
        self.m = m

    def close(self):
        for o in (self.hks, self.ring_ks, self.mult, self.ring):
            if o is not None:
                o.close()

    def evalmult_norelin(self, ct1, ct2):
        out = self.ctx.zeros(3, self.sizeQ, self.N)
        self.mult.evalmult_norelin(self.ctx.upload(ct1), self.ctx.upload(ct2), out)
        return self.ctx.download(out).copy()

    def evalmult(self, ct1, ct2, evk_a, evk_b):
        evk = self.m.EvalKey(self.hks, evk_a, evk_b)
        out = self.ctx.zeros(2, self.sizeQ, self.N)
        self.mult.evalmult(self.hks, self.ctx.upload(ct1), self.ctx.upload(ct2), evk, out)
        res = self.ctx.download(out).copy()
        evk.close()
        return res


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["bfv_n64", "bfv_n128_d4"])
def test_gpu_vs_golden(gpu, golden, name):
    ctx, m = gpu
    g = golden(name)
    d = DeviceBfv(gpu, g)
    try:
        nQ, nR, N = d.sizeQ, d.sizeR, d.N
        o = ctx.zeros(nR, N)
        d.q_to_r.apply(ctx.upload(g["switch_in"]), o)
        assert (ctx.download(o) == g["switch_out"]).all()
        assert (ctx.download(d.expand(ctx.upload(g["expand_in"]))) == g["expand_out"]).all()
        assert (ctx.download(d.fast_expand(ctx.upload(g["fastexp_in"]))) == g["fastexp_out"]).all()
        o = ctx.zeros(nQ, N)
        d.sr.apply(ctx.upload(g["sr_in"]), o)
        assert (ctx.download(o) == g["sr_out"]).all()
        assert (d.evalmult_norelin(g["ct1"], g["ct2"]) == g["evalmult_norelin"]).all()
    finally:
        d.close()
    # the product entry points (ofhe_bfv_evalmult_norelin / ofhe_bfv_evalmult): same vectors
    p = ProductBfv(gpu, g)
    try:
        assert (p.evalmult_norelin(g["ct1"], g["ct2"]) == g["evalmult_norelin"]).all()
        assert (p.evalmult(g["ct1"], g["ct2"], g["evk_a"], g["evk_b"]) == g["evalmult"]).all()
    finally:
        p.close()


@pytest.mark.gpu
def test_gpu_config3_size_vs_oracle(gpu, oracle, golden):
    """BASELINE config 3 shape: N = 2^15, 12 + 12 sixty-bit limbs (log2 q ~ 720).  Tables are the reference's own
    (golden fixture); inputs are seeded, with boundary coefficients (0, q-1, 1) in the first columns."""
    ctx, m = gpu
    g = golden("bfv_cfg3_tables")
    N, nQ, nR = (int(v) for v in g["shape"])
    assert (N, nQ) == (1 << 15, 12)
    d = DeviceBfv(gpu, g)
    T = d.T
    try:
        Q, R = g["moduli"][:nQ], g["moduli"][nQ:]
        xc = seeded_residues(700, Q, N)
        xc[:, 0] = 0
        xc[:, 1] = Q - np.uint64(1)
        xc[:, 2] = 1
        o = ctx.zeros(nR, N)
        d.q_to_r.apply(ctx.upload(xc), o)
        assert (ctx.download(o) == oracle.switch_crt_basis(Q, T["QlHatInvModq"], R, T["QlHatModr"], T["alphaQlModr"],
                                                           T["qInv"], xc)).all()
        assert (ctx.download(d.fast_expand(ctx.upload(xc))) == oracle_fast_expand(oracle, g, T, xc)).all()
        xe = seeded_residues(701, Q, N)
        assert (ctx.download(d.expand(ctx.upload(xe))) == oracle_expand(oracle, g, T, xe)).all()
        xq = seeded_residues(702, g["moduli"], N)
        xq[:, 0] = 0
        xq[:, 1] = g["moduli"] - np.uint64(1)
        o = ctx.zeros(nQ, N)
        d.sr.apply(ctx.upload(xq), o)
        assert (ctx.download(o) == oracle.scale_and_round(Q, True, T["tQlSlHatInvModsDivsModq"],
                                                          T["tQlSlHatInvModsDivsFrac"], xq)).all()
        # the product entry points at this size: the fused device composition equals the step-by-step one (whose
        # pieces are pinned above), and the relinearised product equals NTT + the oracle's hybrid key switch
        ct1 = np.stack([seeded_residues(710, Q, N), seeded_residues(711, Q, N)])
        ct2 = np.stack([seeded_residues(712, Q, N), seeded_residues(713, Q, N)])
        want3 = d.evalmult_norelin(ct1, ct2)
        p = ProductBfv(gpu, g)
        try:
            assert (p.evalmult_norelin(ct1, ct2) == want3).all()
            sizeP, dnum, _ = (int(v) for v in g["ks_shape"])
            qp = np.concatenate([Q, g["moduliP"]])
            ea = seeded_residues(714, np.tile(qp, dnum), N).reshape(dnum, nQ + sizeP, N)
            eb = seeded_residues(715, np.tile(qp, dnum), N).reshape(dnum, nQ + sizeP, N)
            g2 = dict(shape=g["shape"], ks_shape=g["ks_shape"], moduli=g["moduli"], roots=g["roots"],
                      moduliP=g["moduliP"], rootsP=g["rootsP"], evalmult_norelin=want3, evk_a=ea, evk_b=eb)
            assert (p.evalmult(ct1, ct2, ea, eb) == oracle_evalmult_relin(oracle, g2)).all()
        finally:
            p.close()
    finally:
        d.close()


@pytest.mark.gpu
def test_gpu_vs_reference_live(gpu, reference):
    b = reference.bfv(1 << 11, 4, 60, 65537, 0)
    try:
        d = DeviceBfv(gpu, _live_tables(b))
        Q = b.moduli[:b.sizeQ]
        ct1 = np.stack([seeded_residues(1, Q, b.N), seeded_residues(2, Q, b.N)])
        ct2 = np.stack([seeded_residues(3, Q, b.N), seeded_residues(4, Q, b.N)])
        want3 = b.evalmult_norelin(ct1, ct2)
        assert (d.evalmult_norelin(ct1, ct2) == want3).all()
        d.close()
        p = ProductBfv(gpu, _live_tables(b), ks=b.ks_shape())
        try:
            assert (p.evalmult_norelin(ct1, ct2) == want3).all()
            ea, eb = b.evalmult_key()
            assert (p.evalmult(ct1, ct2, ea, eb) == b.evalmult(ct1, ct2)).all()
        finally:
            p.close()
    finally:
        b.close()
