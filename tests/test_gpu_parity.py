"""GPU tier: the CUDA path, called through the C ABI, against the oracle on identical seeded inputs, against the
committed golden vectors from the reference, and (where oracle/_ref travelled) against the reference itself.
Everything is integer arithmetic: the bar is bit-exact equality of every residue."""
import numpy as np
import pytest

from oracle_lib import seeded_residues

pytestmark = pytest.mark.gpu


def _ring(gpu, N, moduli, roots):
    ctx, m = gpu
    return m.Ring(ctx, N, moduli, roots)


def _roundtrip(gpu, ring, x, idx):
    ctx, m = gpu
    d = ctx.upload(x)
    ring.ntt_fwd(d, d, idx)
    y = ctx.download(d).copy()
    ring.ntt_inv(d, d, idx)
    return y, ctx.download(d).copy()


def test_kat_q113(gpu):
    # UTTransform.CRT_polynomial_mult (src/core/unittest/UnitTestTransform.cpp:52-90)
    ctx, m = gpu
    q, N = 113, 4
    root = m.lib().ofhe_root_of_unity(2 * N, q)
    assert root == 18
    ring = m.Ring(ctx, N, [q], [root])
    a = ctx.upload(np.array([[1, 2, 4, 1]], dtype=np.uint64))
    ring.ntt_fwd(a, a, [0])
    ring.binop(2, a, a, a, [0])
    ring.ntt_inv(a, a, [0])
    assert ctx.download(a)[0].tolist() == [94, 109, 11, 18]
    ring.close()


@pytest.mark.parametrize("tag", ["n8", "n64", "n4096"])
def test_ntt_golden(gpu, golden, tag):
    g = golden("core_vectors")
    q, roots = g[f"{tag}_moduli"], g[f"{tag}_roots"]
    N = g[f"{tag}_coeff"].shape[1]
    ring = _ring(gpu, N, q, roots)
    y, back = _roundtrip(gpu, ring, g[f"{tag}_coeff"], list(range(len(q))))
    assert (y == g[f"{tag}_eval"]).all()
    assert (back == g[f"{tag}_coeff"]).all()
    if tag == "n64":
        T, Tp, Ti, Tip, ni, nip = ring.twiddles(0)
        assert (T == g["n64_T"]).all() and (Tp == g["n64_Tp"]).all()
        assert (Ti == g["n64_Ti"]).all() and (Tip == g["n64_Tip"]).all()
        assert [ni, nip] == g["n64_nInv"].tolist()
    ring.close()


@pytest.mark.parametrize("logN", [1, 2, 5, 9, 11, 12, 13, 14, 15, 16, 17])
def test_ntt_vs_oracle(gpu, oracle, logN):
    ctx, m = gpu
    N = 1 << logN
    bits = [60, 59, 45, 30] if logN >= 4 else [28, 20]
    q = np.array([oracle.L.orc_last_prime(b, 2 * N) for b in bits], dtype=np.uint64)
    roots = np.array([oracle.L.orc_root_of_unity(2 * N, int(x)) for x in q], dtype=np.uint64)
    ring = _ring(gpu, N, q, roots)
    x = seeded_residues(1000 + logN, q, N)
    # edge rows: all zero, all q-1 (reference tests use literal small vectors; these stress the lazy ranges)
    x[1] = 0
    x[-1] = q[-1] - np.uint64(1)
    y, back = _roundtrip(gpu, ring, x, list(range(len(q))))
    assert (back == x).all()
    for l in range(len(q)):
        assert (y[l] == oracle.ntt_fwd(q[l], roots[l], x[l])).all(), f"limb {l}"
        assert y[l].max() < q[l]
    # out-of-place + permuted limb->modulus map (metadata-only re-indexing)
    perm = [2, 0] if len(q) > 2 else [1, 0]
    src = ctx.upload(x[perm])
    dst = ctx.zeros(len(perm), N)
    ring.ntt_fwd(src, dst, perm)
    assert (ctx.download(dst) == y[perm]).all()
    assert (ctx.download(src) == x[perm]).all()
    ring.close()


def test_ntt_linearity_and_convolution_full_size(gpu, oracle):
    """Size-independent properties at BASELINE size (N = 2^16): NTT(a)+NTT(b) == NTT(a+b) and the negacyclic
    convolution theorem against a direct O(N) check of one coefficient."""
    ctx, m = gpu
    N = 1 << 16
    q, roots = m.dcrt_params(N, 3, 60)
    ring = _ring(gpu, N, q, roots)
    idx = [0, 1, 2]
    a, b = seeded_residues(1, q, N), seeded_residues(2, q, N)
    da, db = ctx.upload(a), ctx.upload(b)
    ds = ctx.zeros(3, N)
    ring.binop(0, ds, da, db, idx)
    for t in (da, db, ds):
        ring.ntt_fwd(t, t, idx)
    chk = ctx.zeros(3, N)
    ring.binop(0, chk, da, db, idx)
    assert (ctx.download(chk) == ctx.download(ds)).all()
    ring.binop(2, chk, da, db, idx)
    ring.ntt_inv(chk, chk, idx)
    c = ctx.download(chk)
    for l in range(3):
        ql = int(q[l])
        ai, bi = a[l].astype(object), b[l].astype(object)
        for k in (0, 1, 777, N - 1):
            # c_k = sum_{i<=k} a_i b_{k-i} - sum_{i>k} a_i b_{N+k-i}   (X^N = -1)
            pos = int(np.dot(ai[:k + 1], bi[k::-1]))
            neg = int(np.dot(ai[k + 1:], bi[:k:-1])) if k + 1 < N else 0
            assert int(c[l][k]) == (pos - neg) % ql, (l, k)
    ring.close()


def test_limb_ops_golden_and_oracle(gpu, oracle, golden):
    ctx, m = gpu
    g = golden("core_vectors")
    q, roots = g["n64_moduli"], g["n64_roots"]
    ring = _ring(gpu, 64, q, roots)
    a, b = ctx.upload(g["n64_a"]), ctx.upload(g["n64_b"])
    out = ctx.zeros(2, 64)
    for op, name in ((0, "add"), (1, "sub"), (2, "mul"), (3, "neg")):
        ring.binop(op, out, a, b, [0, 1])
        assert (ctx.download(out) == g[f"n64_{name}"]).all(), name
    for k in (5, 127):
        ring.automorphism(out, a, 2, k)
        assert (ctx.download(out) == g[f"n64_auto{k}"]).all()
        assert (ring.automorphism_map(k) == oracle.automorphism_map(64, k)).all()
        # COEFFICIENT-format branch (a15, poly-impl.h:359-362): a zero coefficient with an odd wrap count becomes q
        az = g["n64_a"].copy()
        az[:, :3] = 0
        ring.automorphism_coeff(out, ctx.upload(az), [0, 1], k)
        assert (ctx.download(out) == oracle.dcrt_automorphism_coeff(q, k, az)).all()
    o1 = ctx.zeros(1, 64)
    ring.switch_modulus(o1, a[0:1], 0, 1)
    assert (ctx.download(o1)[0] == g["n64_swmod_down"]).all()
    ring.switch_modulus(o1, b[1:2], 1, 0)
    assert (ctx.download(o1)[0] == g["n64_swmod_up"]).all()
    ring.close()
    # larger, mixed-size moduli, in-place forms, MAC, per-limb scalars
    N = 1 << 13
    qs = np.array([oracle.L.orc_last_prime(bb, 2 * N) for bb in (60, 59, 50, 31)], dtype=np.uint64)
    rs = np.array([oracle.L.orc_root_of_unity(2 * N, int(x)) for x in qs], dtype=np.uint64)
    ring = _ring(gpu, N, qs, rs)
    idx = [0, 1, 2, 3]
    x, y = seeded_residues(5, qs, N), seeded_residues(6, qs, N)
    x[0, :3] = [0, qs[0] - 1, qs[0] - 1]
    y[0, :3] = [0, qs[0] - 1, 1]
    dx, dy = ctx.upload(x), ctx.upload(y)
    for op in range(4):
        o = ctx.zeros(4, N)
        ring.binop(op, o, dx, dy, idx)
        assert (ctx.download(o) == oracle.dcrt_binop(qs, op, x, y)).all(), op
    acc = ctx.upload(x)
    ring.binop(4, acc, dx, dy, idx)  # acc += x*y
    assert (ctx.download(acc) == oracle.dcrt_binop(qs, 0, x, oracle.dcrt_binop(qs, 2, x, y))).all()
    inpl = ctx.upload(x)
    ring.binop(0, inpl, inpl, dy, idx)  # x += y
    assert (ctx.download(inpl) == oracle.dcrt_binop(qs, 0, x, y)).all()
    s = np.array([3, int(qs[1]) - 1, 0, 77777], dtype=np.uint64)
    o = ctx.zeros(4, N)
    ring.mul_scalars(o, dx, s, idx)
    assert (ctx.download(o) == oracle.dcrt_times_scalars(qs, x, s)).all()
    for f, t in [(0, 1), (1, 0), (0, 3), (3, 0), (2, 1)]:
        o1 = ctx.zeros(1, N)
        ring.switch_modulus(o1, dx[f:f + 1], f, t)
        assert (ctx.download(o1)[0] == oracle.switch_modulus(qs[f], qs[t], x[f])).all(), (f, t)
    ring.close()


def test_approx_switch_crt_basis_vs_oracle(gpu, oracle):
    ctx, m = gpu
    N = 1 << 10
    c = oracle.ckks(N, 20, 59, 60, 1, 0)  # one digit of 21 limbs -> 8 aux primes: exercises the > 16 source-limb path
    try:
        ring = _ring(gpu, N, c.moduli, c.roots)
        src = list(range(c.sizeQ))
        dst = list(range(c.sizeQ, c.sizeQP))
        x = seeded_residues(9, c.moduli[:c.sizeQ], N)
        x[:, 0] = 0
        x[:, 1] = c.moduli[:c.sizeQ] - np.uint64(1)
        hatinv = c.PartQlHatInvModq(0, c.sizeQ)
        hatmod, compl = c.PartQlHatModp(c.sizeQ, 0)
        want = oracle.approx_switch_crt_basis(c.moduli[:c.sizeQ], hatinv, compl, hatmod, x)
        d = ctx.upload(x)
        o = ctx.zeros(len(dst), N)
        ring.approx_switch_crt_basis(src, dst, d, o)
        assert (ctx.download(o) == want).all()
        # and the reverse direction P -> first 5 Q limbs (ModDown-shaped)
        P = c.moduli[c.sizeQ:]
        xp = seeded_residues(10, P, N)
        hinv = np.array([pow(int(np.prod([int(v) % int(P[i]) for j, v in enumerate(P) if j != i], dtype=object) % int(P[i])),
                             -1, int(P[i])) for i in range(len(P))], dtype=np.uint64)
        mat = np.array([[int(np.prod([int(v) for j, v in enumerate(P) if j != i], dtype=object) % int(c.moduli[k]))
                         for k in range(5)] for i in range(len(P))], dtype=np.uint64)
        want = oracle.approx_switch_crt_basis(P, hinv, c.moduli[:5], mat, xp)
        d = ctx.upload(xp)
        o = ctx.zeros(5, N)
        ring.approx_switch_crt_basis(dst, list(range(5)), d, o)
        assert (ctx.download(o) == want).all()
        ring.close()
    finally:
        c.close()
    # more than 64 source limbs (the reference's depth-limit tests reach ~140): the run-time-loop kernel
    c = oracle.ckks(1 << 8, 79, 50, 60, 1, 0)  # one digit of 80 limbs -> 27 aux primes
    try:
        N = 1 << 8
        ring = _ring(gpu, N, c.moduli, c.roots)
        x = seeded_residues(11, c.moduli[:c.sizeQ], N)
        x[:, 0] = c.moduli[:c.sizeQ] - np.uint64(1)
        hatmod, compl = c.PartQlHatModp(c.sizeQ, 0)
        want = oracle.approx_switch_crt_basis(c.moduli[:c.sizeQ], c.PartQlHatInvModq(0, c.sizeQ), compl, hatmod, x)
        o = ctx.zeros(c.sizeQP - c.sizeQ, N)
        ring.approx_switch_crt_basis(list(range(c.sizeQ)), list(range(c.sizeQ, c.sizeQP)), ctx.upload(x), o)
        assert (ctx.download(o) == want).all()
        ring.close()
    finally:
        c.close()


def _hks_from_golden(gpu, g):
    ctx, m = gpu
    N, sizeQ, sizeP, dnum, alpha = (int(v) for v in g["shape"])
    ring = m.Ring(ctx, N, g["moduli"], g["roots"])
    hks = m.Hks(ring, sizeQ, sizeP, dnum)
    assert hks.alpha == alpha
    return ring, hks


def _crt_tables(moduli, src, dst):
    """QHatInvModq and QHatModp of the basis {moduli[src]} towards {moduli[dst]} with exact integers
    (dcrtpoly-impl.h:888-932 takes them from CryptoParametersRNS)."""
    q = [int(moduli[i]) for i in src]
    Q = 1
    for v in q:
        Q *= v
    hatinv = np.array([pow((Q // qi) % qi, -1, qi) for qi in q], dtype=np.uint64)
    mat = np.array([[(Q // qi) % int(moduli[j]) for j in dst] for qi in q], dtype=np.uint64)
    return hatinv, mat


@pytest.mark.parametrize("N,ns,nd", [(128, 15, 45), (128, 1, 3), (256, 16, 17), (256, 12, 90), (256, 3, 5),
                                     (4096, 15, 45), (4096, 8, 33), (64, 15, 45)])
def test_approx_switch_crt_basis_tensor_core_shapes(gpu, oracle, N, ns, nd):
    """The tcgen05 base conversion (csrc/baseconv_tc.cu: <= 16 source limbs, N >= 128) over 1..6 accumulator chunks,
    a ragged last chunk, exactly 16 source limbs, a conversion split into two jobs (> 80 targets) and, at N = 64, the
    CUDA-core kernel the library falls back to -- all against the oracle's ApproxSwitchCRTBasis."""
    ctx, m = gpu
    p = m.ckks_paramgen(N, 89, 59, 60, 3, 0)
    mod = p["moduli"]
    ring = _ring(gpu, N, mod, p["roots"])
    src = list(range(2, 2 + ns))
    dst = [i for i in range(len(mod)) if i not in src][:nd]
    x = seeded_residues(31 + ns, mod[src], N)
    x[:, 0] = 0
    x[:, 1] = mod[src] - np.uint64(1)
    hatinv, mat = _crt_tables(mod, src, dst)
    want = oracle.approx_switch_crt_basis(mod[src], hatinv, mod[dst], mat, x)
    o = ctx.zeros(nd, N)
    ring.approx_switch_crt_basis(src, dst, ctx.upload(x), o)
    assert (ctx.download(o) == want).all()
    ring.close()


def test_approx_switch_crt_basis_cuda_core_kernel_still_exact():
    """OFHE_BCONV_TC=0 keeps every conversion on the CUDA-core kernel (baseconv.cu); the switch is read once per
    process, hence the child process.  Same shapes, exact big-integer expectation (tools/bconv_tc_check.py)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OFHE_BCONV_TC="0")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "bconv_tc_check.py"), "parity"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PARITY PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_fused_moddown_tail_variant_still_exact():
    """OFHE_NTT_TAIL=1 (measurement switch, default off) makes the ApproxModDown tail the epilogue of the forward NTT's
    tile pass (ntt_fwd_tile_tail, csrc/ntt.cu).  The switch is read once per process, hence the child: the full-size
    config-2 EvalMult against the oracle with the fused kernel."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OFHE_NTT_TAIL="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_parity.py"), "-q", "-x", "-m", "gpu",
                        "-k", "test_config2_full_size"], env=env, capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["ckks_n64", "ckks_n128_flex"])
def test_hybrid_keyswitch_golden(gpu, golden, name):
    """Every stage of a16-a19 and a14/a15 against vectors produced by the unmodified reference."""
    ctx, m = gpu
    g = golden(name)
    ring, hks = _hks_from_golden(gpu, g)
    N, sizeQ, sizeP = ring.N, hks.sizeQ, hks.sizeP
    # tables (a20) rebuilt from the moduli
    assert (hks.table(0) == g["PInvModq"]).all() and (hks.table(1) == g["PModq"]).all()
    assert (hks.table(2) == g["PHatInvModp"]).all()
    assert (hks.table(3).reshape(sizeP, sizeQ) == g["PHatModq"]).all()
    evk = m.EvalKey(hks, g["evk_a"], g["evk_b"])
    for sizeQl in (sizeQ, sizeQ - 1, 1):
        nd = hks.num_digits(sizeQl)
        for part in range(nd):
            ds = min(hks.alpha, sizeQl - part * hks.alpha)
            assert (hks.table(4, part, ds) == g[f"l{sizeQl}_p{part}_hatinv"]).all()
            assert (hks.table(5, sizeQl, part).reshape(g[f"l{sizeQl}_p{part}_hatmodp"].shape) ==
                    g[f"l{sizeQl}_p{part}_hatmodp"]).all()
        if sizeQl > 1:
            assert (hks.table(6, sizeQl) == g[f"l{sizeQl}_rescaleA"]).all()
            assert (hks.table(7, sizeQl) == g[f"l{sizeQl}_rescaleB"]).all()
        c = ctx.upload(g[f"l{sizeQl}_c"])
        digits = ctx.zeros(nd, sizeQl + sizeP, N)
        hks.modup(sizeQl, c, digits)
        assert (ctx.download(digits) == g[f"l{sizeQl}_digits"]).all(), f"modup l={sizeQl}"
        ks = ctx.zeros(2, sizeQl, N)
        hks.keyswitch_core(sizeQl, c, evk, ks)
        assert (ctx.download(ks) == g[f"l{sizeQl}_ks"]).all(), f"keyswitch l={sizeQl}"
        out = ctx.zeros(2, sizeQl, N)
        hks.evalmult(sizeQl, ctx.upload(g[f"l{sizeQl}_ct1"]), ctx.upload(g[f"l{sizeQl}_ct2"]), evk, out)
        assert (ctx.download(out) == g[f"l{sizeQl}_evalmult"]).all(), f"evalmult l={sizeQl}"
        # same through the host-buffer entry point
        oh = np.zeros((2, sizeQl, N), dtype=np.uint64)
        hks.evalmult_host(sizeQl, np.ascontiguousarray(g[f"l{sizeQl}_ct1"]), np.ascontiguousarray(g[f"l{sizeQl}_ct2"]),
                          evk, oh)
        assert (oh == g[f"l{sizeQl}_evalmult"]).all()
        # and through the software-pipelined batch entry point (more operations than staging slots)
        a1 = np.ascontiguousarray(g[f"l{sizeQl}_ct1"])
        a2 = np.ascontiguousarray(g[f"l{sizeQl}_ct2"])
        outs = [np.zeros((2, sizeQl, N), dtype=np.uint64) for _ in range(7)]
        hks.evalmult_host_batch(sizeQl, [a1, a2] + [a1] * 5, [a2, a1] + [a2] * 5, evk, outs)
        for o in outs:  # the tensor product is symmetric in its two arguments
            assert (o == g[f"l{sizeQl}_evalmult"]).all()
        if sizeQl > 1:
            r = ctx.zeros(2, sizeQl - 1, N)
            hks.rescale(sizeQl, ctx.upload(g[f"l{sizeQl}_ct1"]), r, 2)
            assert (ctx.download(r) == g[f"l{sizeQl}_rescale"]).all(), f"rescale l={sizeQl}"
    rk = m.EvalKey(hks, g["rot_a"], g["rot_b"])
    sizeQl = sizeQ - 1
    k = int(g["rot_k"][0])
    assert k == m.find_automorphism_index(1, 2 * N)
    out = ctx.zeros(2, sizeQl, N)
    hks.rotate(sizeQl, k, ctx.upload(g[f"l{sizeQl}_ct1"]), rk, out)
    assert (ctx.download(out) == g["rot_out"]).all()
    # hoisted rotations against the reference's EvalFastRotation / EvalFastRotationExt / KeySwitchExt / KeySwitchDown
    dct = ctx.upload(g[f"l{sizeQl}_ct1"])
    digits = ctx.zeros(hks.num_digits(sizeQl), sizeQl + sizeP, N)
    hks.modup(sizeQl, dct[1], digits)
    hks.fast_rotation(sizeQl, k, dct, digits, rk, out)
    assert (ctx.download(out) == g["rot_fast"]).all()
    ext = ctx.zeros(2, sizeQl + sizeP, N)
    for add_first in (True, False):
        hks.fast_rotation_ext(sizeQl, k, add_first, dct, digits, rk, ext)
        assert (ctx.download(ext) == g["rot_ext1" if add_first else "rot_ext0"]).all()
        hks.keyswitch_ext(sizeQl, add_first, dct, ext)
        assert (ctx.download(ext) == g["ks_ext1" if add_first else "ks_ext0"]).all()
    hks.keyswitch_down(sizeQl, ctx.upload(g["rot_ext1"]), out)
    assert (ctx.download(out) == g["ks_down"]).all()
    # honest ciphertexts from the reference: the device product is the ciphertext the reference decrypts correctly
    out = ctx.zeros(2, sizeQ, N)
    hks.evalmult(sizeQ, ctx.upload(g["enc1"]), ctx.upload(g["enc2"]), evk, out)
    assert (ctx.download(out) == g["enc_prod"]).all()
    for o in (evk, rk, hks, ring):
        o.close()


@pytest.mark.parametrize("args", [(1 << 12, 8, 50, 60, 3, 0), (1 << 13, 5, 59, 60, 2, 2), (1 << 14, 3, 45, 55, 0, 0)])
def test_hybrid_keyswitch_vs_oracle(gpu, oracle, args):
    """Mid-size contexts (two-kernel NTT path), random keys: stage-by-stage against the oracle."""
    ctx, m = gpu
    N = args[0]
    c = oracle.ckks(*args)
    try:
        p = m.ckks_paramgen(*args)
        assert (p["moduli"] == c.moduli).all() and (p["roots"] == c.roots).all()
        ring = m.Ring(ctx, N, p["moduli"], p["roots"])
        hks = m.Hks(ring, p["sizeQ"], p["sizeP"], p["dnum"])
        ea = seeded_residues(31, np.tile(c.moduli, c.dnum), N).reshape(c.dnum, c.sizeQP, N)
        eb = seeded_residues(32, np.tile(c.moduli, c.dnum), N).reshape(c.dnum, c.sizeQP, N)
        evk = m.EvalKey(hks, ea, eb)
        for sizeQl in sorted({c.sizeQ, c.sizeQ - 1, 2}, reverse=True):
            x = seeded_residues(40 + sizeQl, c.moduli[:sizeQl], N)
            nd = hks.num_digits(sizeQl)
            digits = ctx.zeros(nd, sizeQl + c.sizeP, N)
            hks.modup(sizeQl, ctx.upload(x), digits)
            want = c.modup(sizeQl, x)
            assert (ctx.download(digits) == want).all(), f"modup {sizeQl}"
            e01 = ctx.zeros(2, sizeQl + c.sizeP, N)
            hks.inner_product(sizeQl, digits, evk, e01)
            w0, w1 = c.inner_product(sizeQl, want, ea, eb)
            assert (ctx.download(e01) == np.stack([w0, w1])).all(), f"inner product {sizeQl}"
            md = ctx.zeros(2, sizeQl, N)
            hks.approx_mod_down(sizeQl, e01, md, 2)
            assert (ctx.download(md) == np.stack([c.approx_mod_down(sizeQl, w0), c.approx_mod_down(sizeQl, w1)])).all()
            md1 = ctx.zeros(1, sizeQl, N)
            hks.approx_mod_down(sizeQl, e01[1], md1, 1)
            assert (ctx.download(md1)[0] == c.approx_mod_down(sizeQl, w1)).all()
            ct1 = np.stack([seeded_residues(51, c.moduli[:sizeQl], N), seeded_residues(52, c.moduli[:sizeQl], N)])
            ct2 = np.stack([seeded_residues(53, c.moduli[:sizeQl], N), seeded_residues(54, c.moduli[:sizeQl], N)])
            out = ctx.zeros(2, sizeQl, N)
            hks.evalmult(sizeQl, ctx.upload(ct1), ctx.upload(ct2), evk, out)
            assert (ctx.download(out) == c.evalmult(sizeQl, ct1, ct2, ea, eb)).all(), f"evalmult {sizeQl}"
            r = ctx.zeros(2, sizeQl - 1, N)
            hks.rescale(sizeQl, ctx.upload(ct1), r, 2)
            assert (ctx.download(r) == np.stack([c.rescale_poly(sizeQl, ct1[0]), c.rescale_poly(sizeQl, ct1[1])])).all()
            k = m.find_automorphism_index(3, 2 * N)
            out = ctx.zeros(2, sizeQl, N)
            hks.rotate(sizeQl, k, ctx.upload(ct1), evk, out)
            assert (ctx.download(out) == c.automorphism_ct(sizeQl, k, ct1, ea, eb)).all()
            # hoisted: one ModUp, several rotations (here the same key stands in for every rotation key)
            dct = ctx.upload(ct1)
            hks.modup(sizeQl, dct[1], digits)
            hd = ctx.download(digits)
            ext = ctx.zeros(2, sizeQl + c.sizeP, N)
            for idx in (3, -1):
                kk = m.find_automorphism_index(idx, 2 * N)
                hks.fast_rotation(sizeQl, kk, dct, digits, evk, out)
                assert (ctx.download(out) == c.fast_rotation(sizeQl, kk, ct1, hd, ea, eb)).all(), f"fast rot {idx}"
                hks.fast_rotation_ext(sizeQl, kk, idx == 3, dct, digits, evk, ext)
                want_ext = c.fast_rotation_ext(sizeQl, kk, idx == 3, ct1, hd, ea, eb)
                assert (ctx.download(ext) == want_ext).all(), f"fast rot ext {idx}"
            hks.keyswitch_down(sizeQl, ext, out)
            assert (ctx.download(out) == c.keyswitch_down(sizeQl, want_ext)).all()
            hks.keyswitch_ext(sizeQl, True, dct, ext)
            assert (ctx.download(ext) == c.keyswitch_ext(sizeQl, True, ct1)).all()
        for o in (evk, hks, ring):
            o.close()
    finally:
        c.close()


def test_reference_live_if_present(gpu, reference):
    """Where the compiled reference travelled to this box: one live EvalMult + rotation against it."""
    ctx, m = gpu
    args = (1 << 12, 6, 50, 60, 3, 0)
    rc = reference.ckks(*args)
    try:
        p = m.ckks_paramgen(*args)
        assert (p["moduli"] == rc.moduli).all() and (p["roots"] == rc.roots).all()
        ring = m.Ring(ctx, args[0], p["moduli"], p["roots"])
        hks = m.Hks(ring, p["sizeQ"], p["sizeP"], p["dnum"])
        ea, eb = rc.evalmult_key()
        evk = m.EvalKey(hks, ea, eb)
        sizeQl, N = rc.sizeQ, rc.N
        e1 = rc.encrypt(np.array([1.5, -2.0, 0.25]))
        e2 = rc.encrypt(np.array([2.0, 0.5, -4.0]))
        out = ctx.zeros(2, sizeQl, N)
        hks.evalmult(sizeQl, ctx.upload(e1), ctx.upload(e2), evk, out)
        got = ctx.download(out)
        assert (got == rc.evalmult(sizeQl, e1, e2)).all()
        dec = rc.decrypt(sizeQl, 2, got, 3)
        assert np.allclose(dec, [3.0, -1.0, -1.0], atol=1e-4)
        k, ra, rb = rc.rotate_key(2)
        rk = m.EvalKey(hks, ra, rb)
        hks.rotate(sizeQl, k, ctx.upload(e1), rk, out)
        assert (ctx.download(out) == rc.rotate(sizeQl, 2, e1)).all()
        for o in (evk, rk, hks, ring):
            o.close()
    finally:
        rc.close()


def test_config2_full_size(gpu, oracle):
    """BASELINE config 2 (N = 2^16, depth 44 -> 45 + 15 limbs, dnum 3, alpha 15) at FULL size.
    The oracle (OpenMP C) finishes one EvalMult here in seconds, so the whole pipeline is compared residue for
    residue; on top of that, size-independent properties: NTT round trip on all 45 limbs, ModUp keeps the digit's
    own limbs bit-identical, KeySwitch(0) == 0, determinism."""
    ctx, m = gpu
    N, depth = 1 << 16, 44
    p = m.ckks_paramgen(N, depth, 59, 60, 0, 0)
    assert (p["sizeQ"], p["sizeP"], p["dnum"], p["alpha"]) == (45, 15, 3, 15)
    c = oracle.ckks(N, depth, 59, 60, 0, 0)
    try:
        assert (p["moduli"] == c.moduli).all() and (p["roots"] == c.roots).all()
        ring = m.Ring(ctx, N, p["moduli"], p["roots"])
        hks = m.Hks(ring, 45, 15, 3)
        sizeQl = 45
        idx = list(range(sizeQl))
        ea = seeded_residues(61, np.tile(c.moduli, 3), N).reshape(3, 60, N)
        eb = seeded_residues(62, np.tile(c.moduli, 3), N).reshape(3, 60, N)
        evk = m.EvalKey(hks, ea, eb)
        ct1 = np.stack([seeded_residues(63, c.moduli[:sizeQl], N), seeded_residues(64, c.moduli[:sizeQl], N)])
        ct2 = np.stack([seeded_residues(65, c.moduli[:sizeQl], N), seeded_residues(66, c.moduli[:sizeQl], N)])
        d1, d2 = ctx.upload(ct1), ctx.upload(ct2)
        # NTT round trip on all 45 limbs; two limbs against the oracle
        t = d1[0].clone()
        ring.ntt_inv(t, t, idx)
        coeff = ctx.download(t).copy()
        ring.ntt_fwd(t, t, idx)
        assert (ctx.download(t) == ct1[0]).all()
        for l in (0, 44):
            assert (oracle.ntt_inv(c.moduli[l], c.roots[l], ct1[0][l]) == coeff[l]).all()
        # ModUp: own-digit limbs untouched, everything canonical
        digits = ctx.zeros(3, 60, N)
        hks.modup(sizeQl, d1[1], digits)
        hd = ctx.download(digits)
        for j in range(3):
            assert (hd[j, 15 * j:15 * (j + 1)] == ct1[1][15 * j:15 * (j + 1)]).all()
            assert (hd[j] < c.moduli[:, None]).all()
        assert (hd == c.modup(sizeQl, ct1[1])).all()
        # KeySwitch(0) == 0
        kz = ctx.zeros(2, sizeQl, N)
        hks.keyswitch_core(sizeQl, ctx.zeros(sizeQl, N), evk, kz)
        assert int(ctx.download(kz).max()) == 0
        # the headline operation, device buffers and host buffers, twice (determinism)
        out = ctx.zeros(2, sizeQl, N)
        hks.evalmult(sizeQl, d1, d2, evk, out)
        want = c.evalmult(sizeQl, ct1, ct2, ea, eb)
        assert (ctx.download(out) == want).all()
        oh = np.zeros((2, sizeQl, N), dtype=np.uint64)
        hks.evalmult_host(sizeQl, ct1, ct2, evk, oh)
        assert (oh == want).all()
        # one level down (uneven digits: 44 = 15 + 15 + 14) after a rescale
        r = ctx.zeros(2, sizeQl - 1, N)
        hks.rescale(sizeQl, out, r, 2)
        hr = ctx.download(r)
        assert (hr[0] == c.rescale_poly(sizeQl, want[0])).all() and (hr[1] == c.rescale_poly(sizeQl, want[1])).all()
        out2 = ctx.zeros(2, sizeQl - 1, N)
        hks.evalmult(sizeQl - 1, r, r, evk, out2)
        assert (ctx.download(out2) == c.evalmult(sizeQl - 1, hr, hr, ea, eb)).all()
        # hoisting at full size (size-independent property): the hoisted rotation equals the plain one bit for bit.
        # (KeySwitchDown of the extended-basis rotation is NOT bit-identical to it: the approximate base conversion
        # does not commute with the sign flips of the automorphism; that path is pinned at the oracle's sizes.)
        k = m.find_automorphism_index(5, 2 * N)
        rot, fast = ctx.zeros(2, sizeQl, N), ctx.zeros(2, sizeQl, N)
        hks.rotate(sizeQl, k, d1, evk, rot)
        digits = ctx.zeros(hks.num_digits(sizeQl), sizeQl + c.sizeP, N)
        hks.modup(sizeQl, d1[1], digits)
        hks.fast_rotation(sizeQl, k, d1, digits, evk, fast)
        assert bool((rot == fast).all())
        ext, ext2 = ctx.zeros(2, sizeQl + c.sizeP, N), ctx.zeros(2, sizeQl + c.sizeP, N)
        hks.fast_rotation_ext(sizeQl, k, True, d1, digits, evk, ext)
        hks.fast_rotation_ext(sizeQl, k, False, d1, digits, evk, ext2)
        lift = ctx.zeros(2, sizeQl + c.sizeP, N)
        hks.keyswitch_ext(sizeQl, True, d1, lift)
        # addFirst only adds auto_k(c0 * P mod q_i) to the Q limbs of the first polynomial
        assert bool((ext[1] == ext2[1]).all()) and bool((ext[0, sizeQl:] == ext2[0, sizeQl:]).all())
        assert bool((lift[:, sizeQl:] == 0).all())
        del digits, ext, ext2, lift
        for o in (evk, hks, ring):
            o.close()
    finally:
        c.close()


def test_config4_shape_n17(gpu, oracle):
    """BASELINE config 4 shape (simple-ckks-bootstrapping parameters at ring 2^17: depth 32 -> 33 + 11 limbs, dnum 3,
    SURVEY.md Appendix B row 4a): the building blocks bootstrapping spends its time in -- EvalMult+relin, rescale and a
    hoisted-style rotation -- at N = 2^17 (5-stage strided NTT kernels), full size, against the oracle."""
    ctx, m = gpu
    args = (1 << 17, 32, 59, 60, 0, 0)
    p = m.ckks_paramgen(*args)
    assert (p["sizeQ"], p["sizeP"], p["dnum"], p["alpha"]) == (33, 11, 3, 11)
    N, sizeQ, sizeP, dnum = p["N"], p["sizeQ"], p["sizeP"], p["dnum"]
    c = oracle.ckks(*args)
    try:
        assert (p["moduli"] == c.moduli).all() and (p["roots"] == c.roots).all()
        ring = m.Ring(ctx, N, p["moduli"], p["roots"])
        hks = m.Hks(ring, sizeQ, sizeP, dnum)
        ea = seeded_residues(91, np.tile(c.moduli, dnum), N).reshape(dnum, sizeQ + sizeP, N)
        eb = seeded_residues(92, np.tile(c.moduli, dnum), N).reshape(dnum, sizeQ + sizeP, N)
        evk = m.EvalKey(hks, ea, eb)
        sizeQl = sizeQ
        ct1 = np.stack([seeded_residues(93, c.moduli[:sizeQl], N), seeded_residues(94, c.moduli[:sizeQl], N)])
        ct2 = np.stack([seeded_residues(95, c.moduli[:sizeQl], N), seeded_residues(96, c.moduli[:sizeQl], N)])
        out = ctx.zeros(2, sizeQl, N)
        hks.evalmult(sizeQl, ctx.upload(ct1), ctx.upload(ct2), evk, out)
        want = c.evalmult(sizeQl, ct1, ct2, ea, eb)
        assert (ctx.download(out) == want).all()
        r = ctx.zeros(2, sizeQl - 1, N)
        hks.rescale(sizeQl, out, r, 2)
        assert (ctx.download(r)[1] == c.rescale_poly(sizeQl, want[1])).all()
        k = m.find_automorphism_index(-5, 2 * N)
        rot = ctx.zeros(2, sizeQl, N)
        hks.rotate(sizeQl, k, ctx.upload(ct1), evk, rot)
        assert (ctx.download(rot) == c.automorphism_ct(sizeQl, k, ct1, ea, eb)).all()
        for o in (evk, hks, ring):
            o.close()
    finally:
        c.close()


@pytest.mark.gpu
def test_upload_cereal_archive(gpu):
    """f3: archive bytes of the reference's wire format -> device rows of a ring (ofhe_dcrtpoly_upload_cereal); the
    archive is untrusted: wrong ring dimension, wrong modulus, too many towers or a non-canonical residue are rejected"""
    ctx, m = gpu
    N, L = 1 << 12, 5
    q, roots = m.dcrt_params(N, L, 60)
    data = seeded_residues(31, q, N)
    blob = m.cereal_write_dcrtpoly(q, roots, 0, data)  # byte-identical to Serial::Serialize (tests/test_cereal_io.py)
    ring = m.Ring(ctx, N, q, roots)
    slab, info = ring.upload_cereal(blob)
    assert (info["N"], info["n_towers"], info["format"]) == (N, L, 0)
    assert (ctx.download(slab) == data).all()
    # the uploaded rows are usable as they are: NTT round trip on the device
    tmp = ctx.zeros(L, N)
    ring.ntt_inv(slab, tmp, list(range(L)))
    ring.ntt_fwd(tmp, tmp, list(range(L)))
    assert (ctx.download(tmp) == data).all()
    # towers in another order of the ring
    perm = [4, 0, 3, 1, 2]
    slab2, _ = m.Ring(ctx, N, q[perm], roots[perm]).upload_cereal(m.cereal_write_dcrtpoly(q, roots, 0, data),
                                                                     mod_idx=[perm.index(i) for i in range(L)])
    assert (ctx.download(slab2) == data).all()
    # rejections
    q2, r2 = m.dcrt_params(2 * N, L, 60)
    with pytest.raises(RuntimeError, match="ring dimension"):
        ring.upload_cereal(m.cereal_write_dcrtpoly(q2, r2, 0, seeded_residues(32, q2, 2 * N)))
    with pytest.raises(RuntimeError, match="modulus differs"):
        ring.upload_cereal(m.cereal_write_dcrtpoly(q[::-1].copy(), roots[::-1].copy(), 0, data[::-1].copy()))
    with pytest.raises(RuntimeError, match="more towers"):
        m.Ring(ctx, N, q[:3], roots[:3]).upload_cereal(blob)
    bad = data.copy()
    bad[2, 17] = q[2]
    with pytest.raises(RuntimeError, match="residue"):
        ring.upload_cereal(m.cereal_write_dcrtpoly(q, roots, 0, bad))
    slab3, _ = ring.upload_cereal(m.cereal_write_dcrtpoly(q, roots, 0, bad), check_residues=False)
    assert (ctx.download(slab3) == bad).all()


@pytest.mark.gpu
@pytest.mark.parametrize("args,lvl_drop", [((1 << 10, 5, 45, 55, 3, 1), 0), ((1 << 12, 4, 50, 60, 2, 0), 1)])
def test_bsgs_linear_transform_level(gpu, oracle, args, lvl_drop):
    """ofhe_ckks_bsgs_level (double-hoisted baby-step / giant-step level, ckksrns-fhe.cpp:1934-1987) with
    every intermediate on the device, bit-exact against the oracle's composition (itself pinned to the reference's own
    primitives run in that order, tests/test_oracle_vs_reference.py::test_bsgs_linear_transform_level)."""
    ctx, m = gpu
    N = args[0]
    c = oracle.ckks(*args)
    try:
        p = m.ckks_paramgen(*args)
        ring = m.Ring(ctx, N, p["moduli"], p["roots"])
        hks = m.Hks(ring, p["sizeQ"], p["sizeP"], p["dnum"])
        sizeQl = c.sizeQ - lvl_drop
        rot_in, rot_out = [-1, 0, 1, 2], [0, 4, 0, 8]
        g, b = len(rot_in), len(rot_out)
        keymod = np.tile(c.moduli, c.dnum)
        host_keys, dev_keys = {}, {}
        for n, idx in enumerate(sorted(set(rot_in + rot_out) - {0})):
            ka = seeded_residues(200 + 2 * n, keymod, N).reshape(c.dnum, c.sizeQP, N)
            kb = seeded_residues(201 + 2 * n, keymod, N).reshape(c.dnum, c.sizeQP, N)
            host_keys[idx] = (ka, kb)
            dev_keys[idx] = m.EvalKey(hks, ka, kb)
        k_of = lambda i: m.find_automorphism_index(i, 2 * N) if i else 0  # noqa: E731
        k_in, k_out = [k_of(i) for i in rot_in], [k_of(i) for i in rot_out]
        qlp = c.moduli_QlP(sizeQl)
        A = np.stack([seeded_residues(100 + d, qlp, N) for d in range(b * g)])
        ct = np.stack([seeded_residues(31, c.moduli[:sizeQl], N), seeded_residues(32, c.moduli[:sizeQl], N)])
        skip = b * g - 2
        want = c.bsgs_level(sizeQl, k_in, k_out, [host_keys.get(i) for i in rot_in], [host_keys.get(i) for i in rot_out],
                            A, skip, ct, oracle.dcrt_binop)
        out = ctx.zeros(2, sizeQl, N)
        ws = hks.bsgs_workspace(sizeQl, g)
        hks.bsgs_level(sizeQl, k_in, k_out, [dev_keys.get(i) for i in rot_in], [dev_keys.get(i) for i in rot_out],
                       ctx.upload(A), skip, ctx.upload(ct), out, ws)
        assert (ctx.download(out) == want).all()
        for o in list(dev_keys.values()) + [hks, ring]:
            o.close()
    finally:
        c.close()
