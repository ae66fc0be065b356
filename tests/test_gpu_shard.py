"""Limb-sharded EvalMult (SURVEY.md 8 e): G virtual ranks on ONE GPU (each with its own context / scratch, the
all-gathers emulated by device copies) must reproduce the unsharded device result -- which test_gpu_parity.py pins to
the oracle / reference -- bit for bit.  The real NCCL path is tests/test_gpu_multi.py (needs >= 2 GPUs)."""
import numpy as np
import pytest

from oracle_lib import seeded_residues

pytestmark = pytest.mark.gpu


def run_virtual(m, ctx0, p, world, sizeQl, ct1, ct2, ea, eb, push=False):
    """Returns the full [2][sizeQl][N] result assembled from the ranks' local rows."""
    import torch

    N, sizeQ, sizeP, dnum = p["N"], p["sizeQ"], p["sizeP"], p["dnum"]
    ranks = []
    for r in range(world):
        ctx = m.Context(0)
        ring = m.Ring(ctx, N, p["moduli"], p["roots"])
        sh = m.Shard(ring, sizeQ, sizeP, dnum, r, world)
        info = sh.info(sizeQl)
        own_q, own_p = sh.owned(sizeQl)
        plan = m.shard_plan(sizeQ, sizeP, sizeQl, world, r)
        assert own_q == plan["own_q"] and own_p == plan["own_p"]
        assert (info["slots_q"], info["slots_p"], info["n_key_rows"]) == (plan["slots_q"], plan["slots_p"], plan["n_key_rows"])
        key_rows = plan["own_q_full"] + [sizeQ + j for j in own_p]
        d = dict(ctx=ctx, ring=ring, sh=sh, info=info, own_q=own_q,
                 ct1=ctx.upload(ct1[:, own_q]) if own_q else ctx.zeros(2, 0, N),
                 ct2=ctx.upload(ct2[:, own_q]) if own_q else ctx.zeros(2, 0, N),
                 out=ctx.zeros(2, len(own_q), N),
                 ea=ctx.upload(ea[:, key_rows]), eb=ctx.upload(eb[:, key_rows]),
                 gq=ctx.zeros(world, info["slots_q"], N), gp=ctx.zeros(world, 2, info["slots_p"], N))
        ranks.append(d)

    def exchange(key):
        torch.cuda.synchronize()
        for src in range(world):
            for dst in range(world):
                if dst != src:
                    ranks[dst][key][src].copy_(ranks[src][key][src])
        torch.cuda.synchronize()

    if push:
        # exchange by peer stores + device barriers: every virtual rank on its own stream, no host synchronisation
        # between the phases; run twice back to back on the same buffers (reuse safety)
        nfl = m.lib().ofhe_peer_flags_words(world, 2)
        flags = [ctx0.zeros(nfl) for _ in range(world)]
        gq_arr = m.ptr_array([d["gq"] for d in ranks])
        gp_arr = m.ptr_array([d["gp"] for d in ranks])
        fl_arr = m.ptr_array(flags)
        streams = [torch.cuda.Stream() for _ in range(world)]
        torch.cuda.synchronize()
        for _ in range(2):
            for r, d in enumerate(ranks):
                with torch.cuda.stream(streams[r]):
                    d["sh"].evalmult_push(sizeQl, d["ct1"], d["ct2"], d["ea"], d["eb"], d["out"], gq_arr, gp_arr, fl_arr)
        torch.cuda.synchronize()
        for r in range(world):
            assert ctx0.peer_barrier_timeouts(flags[r].data_ptr(), world, 2) == 0, "device barrier timed out"
    else:
        for d in ranks:
            d["sh"].phase_a(sizeQl, d["ct1"], d["ct2"], d["out"], d["gq"])
        exchange("gq")
        for d in ranks:
            d["sh"].phase_b(sizeQl, d["gq"], d["ea"], d["eb"], d["gp"])
        exchange("gp")
        for d in ranks:
            d["sh"].phase_c(sizeQl, d["gp"], d["out"])
        torch.cuda.synchronize()
    full = np.zeros((2, sizeQl, N), dtype=np.uint64)
    for d in ranks:
        if d["own_q"]:
            full[:, d["own_q"]] = ctx0.download(d["out"])
    for d in ranks:
        d["sh"].close()
        d["ring"].close()
        d["ctx"].close()
    return full


@pytest.mark.parametrize("args,worlds", [((1 << 12, 8, 50, 60, 3, 0), (1, 2, 3, 8)), ((1 << 13, 5, 59, 60, 2, 0), (2, 4)),
                                         ((1 << 12, 2, 45, 55, 0, 0), (2, 8))])
def test_virtual_ranks_match_unsharded(gpu, oracle, args, worlds):
    ctx, m = gpu
    p = m.ckks_paramgen(*args)
    N, sizeQ, sizeP, dnum = p["N"], p["sizeQ"], p["sizeP"], p["dnum"]
    ring = m.Ring(ctx, N, p["moduli"], p["roots"])
    hks = m.Hks(ring, sizeQ, sizeP, dnum)
    ea = seeded_residues(81, np.tile(p["moduli"], dnum), N).reshape(dnum, sizeQ + sizeP, N)
    eb = seeded_residues(82, np.tile(p["moduli"], dnum), N).reshape(dnum, sizeQ + sizeP, N)
    evk = m.EvalKey(hks, ea, eb)
    for sizeQl in sorted({sizeQ, max(1, sizeQ - 1), 1}, reverse=True):
        ct1 = np.stack([seeded_residues(83, p["moduli"][:sizeQl], N), seeded_residues(84, p["moduli"][:sizeQl], N)])
        ct2 = np.stack([seeded_residues(85, p["moduli"][:sizeQl], N), seeded_residues(86, p["moduli"][:sizeQl], N)])
        out = ctx.zeros(2, sizeQl, N)
        hks.evalmult(sizeQl, ctx.upload(ct1), ctx.upload(ct2), evk, out)
        want = ctx.download(out).copy()
        if sizeQl == sizeQ:  # anchor the unsharded result to the oracle once per context
            c = oracle.ckks(*args)
            try:
                assert (want == c.evalmult(sizeQl, ct1, ct2, ea, eb)).all()
            finally:
                c.close()
        for world in worlds:
            got = run_virtual(m, ctx, p, world, sizeQl, ct1, ct2, ea, eb)
            assert (got == want).all(), f"world={world} sizeQl={sizeQl}"
            if world > 1:
                got = run_virtual(m, ctx, p, world, sizeQl, ct1, ct2, ea, eb, push=True)
                assert (got == want).all(), f"peer-push world={world} sizeQl={sizeQl}"
    for o in (evk, hks, ring):
        o.close()


def run_virtual_rotate_rescale(m, ctx0, p, world, sizeQl, k, ct, ea, eb, push):
    """Limb-sharded EvalAutomorphism(k) followed by a limb-sharded rescale of the same input, on `world` virtual
    ranks of one GPU; returns (rotated [2][sizeQl][N], rescaled [2][sizeQl-1][N]) assembled from the ranks' rows."""
    import torch

    N, sizeQ, sizeP, dnum = p["N"], p["sizeQ"], p["sizeP"], p["dnum"]
    ranks = []
    for r in range(world):
        ctx = m.Context(0)
        ring = m.Ring(ctx, N, p["moduli"], p["roots"])
        sh = m.Shard(ring, sizeQ, sizeP, dnum, r, world)
        info = sh.info(sizeQl)
        own_q, own_p = sh.owned(sizeQl)
        plan = m.shard_plan(sizeQ, sizeP, sizeQl, world, r)
        key_rows = plan["own_q_full"] + [sizeQ + j for j in own_p]
        low = [i for i in own_q if i != sizeQl - 1]
        ranks.append(dict(ctx=ctx, ring=ring, sh=sh, own_q=own_q, low=low,
                          ct=ctx.upload(ct[:, own_q]) if own_q else ctx.zeros(2, 0, N),
                          tmp=ctx.zeros(2, max(len(own_q), 1), N), out=ctx.zeros(2, max(len(own_q), 1), N),
                          res=ctx.zeros(2, max(len(low), 1), N),
                          ea=ctx.upload(ea[:, key_rows]), eb=ctx.upload(eb[:, key_rows]),
                          gq=ctx.zeros(world, info["slots_q"], N), gp=ctx.zeros(world, 2, info["slots_p"], N),
                          gl=ctx.zeros(2, N)))

    if push:
        nfl = m.lib().ofhe_peer_flags_words(world, 3)
        flags = [ctx0.zeros(nfl) for _ in range(world)]
        gq_arr = m.ptr_array([d["gq"] for d in ranks])
        gp_arr = m.ptr_array([d["gp"] for d in ranks])
        gl_arr = m.ptr_array([d["gl"] for d in ranks])
        fl_arr = m.ptr_array(flags)
        streams = [torch.cuda.Stream() for _ in range(world)]
        torch.cuda.synchronize()
        for r, d in enumerate(ranks):
            with torch.cuda.stream(streams[r]):
                d["sh"].rotate_push(sizeQl, k, d["ct"], d["ea"], d["eb"], d["tmp"], d["out"], gq_arr, gp_arr, fl_arr)
                if sizeQl >= 2:
                    d["sh"].rescale_push(sizeQl, d["ct"], d["res"], gl_arr, fl_arr, index=2)
        torch.cuda.synchronize()
        for r in range(world):
            assert ctx0.peer_barrier_timeouts(flags[r].data_ptr(), world, 3) == 0, "device barrier timed out"
    else:
        L = m.lib()

        def exchange(key):
            torch.cuda.synchronize()
            for src in range(world):
                for dst in range(world):
                    if dst != src:
                        ranks[dst][key][src].copy_(ranks[src][key][src])
            torch.cuda.synchronize()

        for d in ranks:
            m._ck(L.ofhe_shard_rotate_a(d["sh"].h, sizeQl, m._ptr(d["ct"]), m._ptr(d["tmp"]), m._ptr(d["gq"]), m._stream()))
        exchange("gq")
        for d in ranks:
            d["sh"].phase_b(sizeQl, d["gq"], d["ea"], d["eb"], d["gp"])
        exchange("gp")
        for d in ranks:
            m._ck(L.ofhe_shard_rotate_c(d["sh"].h, sizeQl, k, m._ptr(d["gp"]), m._ptr(d["tmp"]), m._ptr(d["out"]), m._stream()))
        if sizeQl >= 2:
            owner = (sizeQl - 1) % world
            for d in ranks:
                m._ck(L.ofhe_shard_rescale_a(d["sh"].h, sizeQl, m._ptr(d["ct"]), m._ptr(d["gl"]), m._stream()))
            torch.cuda.synchronize()
            for r, d in enumerate(ranks):  # the caller's broadcast from the owner of the dropped limb
                if r != owner:
                    d["gl"].copy_(ranks[owner]["gl"])
            torch.cuda.synchronize()
            for d in ranks:
                m._ck(L.ofhe_shard_rescale_b(d["sh"].h, sizeQl, m._ptr(d["ct"]), m._ptr(d["gl"]), m._ptr(d["res"]), m._stream()))
        torch.cuda.synchronize()
    rot = np.zeros((2, sizeQl, N), dtype=np.uint64)
    res = np.zeros((2, max(sizeQl - 1, 0), N), dtype=np.uint64)
    for d in ranks:
        if d["own_q"]:
            rot[:, d["own_q"]] = ctx0.download(d["out"])[:, :len(d["own_q"])]
        if d["low"] and sizeQl >= 2:
            res[:, d["low"]] = ctx0.download(d["res"])[:, :len(d["low"])]
    for d in ranks:
        d["sh"].close()
        d["ring"].close()
        d["ctx"].close()
    return rot, res


@pytest.mark.parametrize("args,worlds", [((1 << 12, 8, 50, 60, 3, 0), (2, 3, 8)), ((1 << 13, 5, 59, 60, 2, 0), (4,))])
def test_virtual_ranks_rotate_and_rescale_match_unsharded(gpu, args, worlds):
    """The two other operations a bootstrapping-sized circuit needs, limb-sharded (csrc/shard.cu: rotate_a / evalmult_b /
    rotate_c, rescale_a / rescale_b): bit-identical to the unsharded device entry points (ofhe_ckks_rotate /
    ofhe_ckks_rescale, themselves pinned to the oracle and the reference in test_gpu_parity.py), with the exchange done
    by host copies and by peer stores + device barriers."""
    ctx, m = gpu
    p = m.ckks_paramgen(*args)
    N, sizeQ, sizeP, dnum = p["N"], p["sizeQ"], p["sizeP"], p["dnum"]
    ring = m.Ring(ctx, N, p["moduli"], p["roots"])
    hks = m.Hks(ring, sizeQ, sizeP, dnum)
    ea = seeded_residues(91, np.tile(p["moduli"], dnum), N).reshape(dnum, sizeQ + sizeP, N)
    eb = seeded_residues(92, np.tile(p["moduli"], dnum), N).reshape(dnum, sizeQ + sizeP, N)
    evk = m.EvalKey(hks, ea, eb)
    k = m.find_automorphism_index(3, 2 * N)
    for sizeQl in sorted({sizeQ, max(2, sizeQ - 2)}, reverse=True):
        ct = np.stack([seeded_residues(93, p["moduli"][:sizeQl], N), seeded_residues(94, p["moduli"][:sizeQl], N)])
        d = ctx.upload(ct)
        o = ctx.zeros(2, sizeQl, N)
        hks.rotate(sizeQl, k, d, evk, o)
        want_rot = ctx.download(o).copy()
        o2 = ctx.zeros(2, sizeQl - 1, N)
        hks.rescale(sizeQl, d, o2)
        want_res = ctx.download(o2).copy()
        for world in worlds:
            for push in (False, True):
                rot, res = run_virtual_rotate_rescale(m, ctx, p, world, sizeQl, k, ct, ea, eb, push)
                assert (rot == want_rot).all(), f"rotate world={world} sizeQl={sizeQl} push={push}"
                assert (res == want_res).all(), f"rescale world={world} sizeQl={sizeQl} push={push}"
    for o_ in (evk, hks, ring):
        o_.close()
