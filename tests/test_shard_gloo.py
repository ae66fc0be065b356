"""CPU tier, world_size 2 over gloo: the limb-sharded EvalMult decomposition (ownership plan from
ofhe_b200.shard_plan + two all-gathers) executed with the ORACLE's arithmetic on each rank's limbs must reproduce the
oracle's unsharded EvalMult bit for bit.  This covers the N > 1 host logic (plan, gather layouts, key sharding) without
a GPU; the same plan drives the CUDA path (tests/test_gpu_shard.py, tests/test_gpu_multi.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = (1 << 8, 5, 45, 55, 3, 0)


def _rank_main(rank, world, port, sizeQl, result_q):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ofhe_b200 as m
    from oracle_lib import Oracle, seeded_residues

    o = Oracle()
    c = o.ckks(*ARGS)
    N, sizeQ, sizeP, dnum, alpha = c.N, c.sizeQ, c.sizeP, c.dnum, c.alpha
    q, roots = c.moduli, c.roots
    plan = m.shard_plan(sizeQ, sizeP, sizeQl, world, rank)
    own_q, own_p = plan["own_q"], plan["own_p"]
    ea = seeded_residues(1, np.tile(q, dnum), N).reshape(dnum, sizeQ + sizeP, N)
    eb = seeded_residues(2, np.tile(q, dnum), N).reshape(dnum, sizeQ + sizeP, N)
    ct1 = np.stack([seeded_residues(3, q[:sizeQl], N), seeded_residues(4, q[:sizeQl], N)])
    ct2 = np.stack([seeded_residues(5, q[:sizeQl], N), seeded_residues(6, q[:sizeQl], N)])
    # every rank only touches its own rows from here on
    a0, a1, b0, b1 = ct1[0, own_q], ct1[1, own_q], ct2[0, own_q], ct2[1, own_q]
    mq = q[own_q]
    mul = lambda x, y: o.dcrt_binop(mq, 2, x, y)
    add = lambda x, y: o.dcrt_binop(mq, 0, x, y)
    c0, c1, c2 = mul(a0, b0), add(mul(a0, b1), mul(a1, b0)), mul(a1, b1)

    def all_gather(buf):  # in place on a [world, ...] array, like the NCCL call of the product
        t = torch.from_numpy(buf.view(np.int64))
        outs = [torch.empty_like(t[0]) for _ in range(world)]
        dist.all_gather(outs, t[rank].clone())
        for r in range(world):
            t[r].copy_(outs[r])

    # phase A: INTT of the owned limbs of c2 into this rank's rows of gatherQ
    gq = np.zeros((world, plan["slots_q"], N), dtype=np.uint64)
    for t, i in enumerate(own_q):
        gq[rank, t] = o.ntt_inv(q[i], roots[i], c2[t])
    all_gather(gq)
    gqf = gq.reshape(world * plan["slots_q"], N)
    # phase B: ModUp into the owned rows of every digit, NTT, inner product with the owned key rows
    nd = c.num_digits(sizeQl)
    own_qp = own_q + [sizeQ + j for j in own_p]  # modulus index of each local QlP row
    e0 = np.zeros((len(own_qp), N), dtype=np.uint64)
    e1 = np.zeros_like(e0)
    for j in range(nd):
        s, ds = j * alpha, c.digit_size(sizeQl, j)
        src = np.stack([gqf[plan["gather_q_row"](i)] for i in range(s, s + ds)])
        hatinv = c.PartQlHatInvModq(j, ds)
        hatmod, compl = c.PartQlHatModp(sizeQl, j)
        compl = compl.tolist()
        D = np.zeros((len(own_qp), N), dtype=np.uint64)
        tgt_rows = [t for t, mi in enumerate(own_qp) if not (mi < sizeQ and s <= mi < s + ds)]
        if tgt_rows:
            cols = [compl.index(int(q[own_qp[t]])) for t in tgt_rows]
            conv = o.approx_switch_crt_basis(q[s:s + ds], hatinv, q[[own_qp[t] for t in tgt_rows]],
                                             np.ascontiguousarray(hatmod[:, cols]), src)
            for k, t in enumerate(tgt_rows):
                D[t] = o.ntt_fwd(q[own_qp[t]], roots[own_qp[t]], conv[k])
        for t, i in enumerate(own_q):
            if s <= i < s + ds:
                D[t] = c2[t]
        for t, mi in enumerate(own_qp):
            one = np.array([q[mi]], dtype=np.uint64)
            e0[t] = o.dcrt_binop(one, 0, e0[t:t + 1], o.dcrt_binop(one, 2, D[t:t + 1], eb[j, mi][None]))[0]
            e1[t] = o.dcrt_binop(one, 0, e1[t:t + 1], o.dcrt_binop(one, 2, D[t:t + 1], ea[j, mi][None]))[0]
    gp = np.zeros((world, 2, plan["slots_p"], N), dtype=np.uint64)
    for t, j in enumerate(own_p):
        mi = sizeQ + j
        gp[rank, 0, t] = o.ntt_inv(q[mi], roots[mi], e0[len(own_q) + t])
        gp[rank, 1, t] = o.ntt_inv(q[mi], roots[mi], e1[len(own_q) + t])
    all_gather(gp)
    gpf = gp.reshape(world * 2 * plan["slots_p"], N)
    # phase C: ModDown of the owned Q rows, final additions
    PHatInv, PHat, PInv = c.PHatInvModp(), c.PHatModq(), c.PInvModq()
    out = np.zeros((2, len(own_q), N), dtype=np.uint64)
    for b, (e, cc) in enumerate(((e0, c0), (e1, c1))):
        if not own_q:
            continue
        src = np.stack([gpf[plan["gather_p_row"](j, b)] for j in range(sizeP)])
        conv = o.approx_switch_crt_basis(q[sizeQ:], PHatInv, mq, np.ascontiguousarray(PHat[:, own_q]), src)
        for t, i in enumerate(own_q):
            cv = o.ntt_fwd(q[i], roots[i], conv[t])
            one = np.array([q[i]], dtype=np.uint64)
            d = o.dcrt_binop(one, 1, e[t:t + 1], cv[None])
            d = o.dcrt_times_scalars(one, d, PInv[i:i + 1])
            out[b, t] = o.dcrt_binop(one, 0, d, cc[t:t + 1])[0]
    # collect on rank 0 and compare with the unsharded oracle
    pad = np.zeros((2, plan["slots_q"], N), dtype=np.int64)
    pad[:, :len(own_q)] = out.view(np.int64)
    outs = [torch.zeros(2, plan["slots_q"], N, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(outs, torch.from_numpy(pad))
    if rank == 0:
        full = np.zeros((2, sizeQl, N), dtype=np.uint64)
        for r in range(world):
            rows = [i for i in range(sizeQl) if i % world == r]
            full[:, rows] = outs[r].numpy().view(np.uint64)[:, :len(rows)]
        want = c.evalmult(sizeQl, ct1, ct2, ea, eb)
        result_q.put(bool((full == want).all()))
    c.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,sizeQl", [(2, 6), (2, 5), (3, 6)])
def test_sharded_decomposition_over_gloo(world, sizeQl):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 10 + sizeQl
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, sizeQl, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def _rank_rescale_rotate(rank, world, port, sizeQl, k_index, result_q):
    """Limb-sharded rescale (one broadcast of the dropped limb's coefficient form) and limb-sharded EvalAutomorphism
    (the EvalMult key switch applied to c1, then a limb-local permutation), each rank with the ORACLE's arithmetic on its
    own limbs only: csrc/shard.cu rescale_a/_b and rotate_a / evalmult_b / rotate_c follow exactly this plan."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ofhe_b200 as m
    from oracle_lib import Oracle, seeded_residues

    o = Oracle()
    c = o.ckks(*ARGS)
    N, sizeQ, sizeP, dnum, alpha = c.N, c.sizeQ, c.sizeP, c.dnum, c.alpha
    q, roots = c.moduli, c.roots
    plan = m.shard_plan(sizeQ, sizeP, sizeQl, world, rank)
    own_q, own_p = plan["own_q"], plan["own_p"]
    ea = seeded_residues(11, np.tile(q, dnum), N).reshape(dnum, sizeQ + sizeP, N)
    eb = seeded_residues(12, np.tile(q, dnum), N).reshape(dnum, sizeQ + sizeP, N)
    ct = np.stack([seeded_residues(13, q[:sizeQl], N), seeded_residues(14, q[:sizeQl], N)])
    k = m.find_automorphism_index(k_index, 2 * N)
    l = sizeQl - 1
    # ---- rescale: the owner of limb l broadcasts INTT(c_b[l]); every rank finishes its own limbs below l
    A, B = c.rescale_tables(sizeQl)
    z = np.zeros((2, N), dtype=np.uint64)
    if l % world == rank:
        for b in range(2):
            z[b] = o.ntt_inv(q[l], roots[l], ct[b, l])
    t = torch.from_numpy(z.view(np.int64))
    dist.broadcast(t, src=l % world)
    low = [i for i in own_q if i != l]
    res = np.zeros((2, plan["slots_q"], N), dtype=np.uint64)
    for b in range(2):
        for tt, i in enumerate(low):
            one = np.array([q[i]], dtype=np.uint64)
            sw = o.switch_modulus(q[l], q[i], z[b])[None]
            sw = o.ntt_fwd(q[i], roots[i], o.dcrt_times_scalars(one, sw, A[i:i + 1])[0])[None]
            res[b, tt] = o.dcrt_binop(one, 0, o.dcrt_times_scalars(one, ct[b, i][None], B[i:i + 1]), sw)[0]
    # ---- rotation: key switch of c1 on the sharded plan (phases as in _rank_main), addend (c0 | 0), then the permutation
    def all_gather(buf):
        tt_ = torch.from_numpy(buf.view(np.int64))
        outs = [torch.empty_like(tt_[0]) for _ in range(world)]
        dist.all_gather(outs, tt_[rank].clone())
        for r in range(world):
            tt_[r].copy_(outs[r])

    c1 = ct[1, own_q]
    gq = np.zeros((world, plan["slots_q"], N), dtype=np.uint64)
    for tt, i in enumerate(own_q):
        gq[rank, tt] = o.ntt_inv(q[i], roots[i], c1[tt])
    all_gather(gq)
    gqf = gq.reshape(world * plan["slots_q"], N)
    nd = c.num_digits(sizeQl)
    own_qp = own_q + [sizeQ + j for j in own_p]
    e0 = np.zeros((len(own_qp), N), dtype=np.uint64)
    e1 = np.zeros_like(e0)
    for j in range(nd):
        s, ds = j * alpha, c.digit_size(sizeQl, j)
        src = np.stack([gqf[plan["gather_q_row"](i)] for i in range(s, s + ds)])
        hatinv = c.PartQlHatInvModq(j, ds)
        hatmod, compl = c.PartQlHatModp(sizeQl, j)
        compl = compl.tolist()
        D = np.zeros((len(own_qp), N), dtype=np.uint64)
        tgt_rows = [tt for tt, mi in enumerate(own_qp) if not (mi < sizeQ and s <= mi < s + ds)]
        if tgt_rows:
            cols = [compl.index(int(q[own_qp[tt]])) for tt in tgt_rows]
            conv = o.approx_switch_crt_basis(q[s:s + ds], hatinv, q[[own_qp[tt] for tt in tgt_rows]],
                                             np.ascontiguousarray(hatmod[:, cols]), src)
            for kk, tt in enumerate(tgt_rows):
                D[tt] = o.ntt_fwd(q[own_qp[tt]], roots[own_qp[tt]], conv[kk])
        for tt, i in enumerate(own_q):
            if s <= i < s + ds:
                D[tt] = c1[tt]
        for tt, mi in enumerate(own_qp):
            one = np.array([q[mi]], dtype=np.uint64)
            e0[tt] = o.dcrt_binop(one, 0, e0[tt:tt + 1], o.dcrt_binop(one, 2, D[tt:tt + 1], eb[j, mi][None]))[0]
            e1[tt] = o.dcrt_binop(one, 0, e1[tt:tt + 1], o.dcrt_binop(one, 2, D[tt:tt + 1], ea[j, mi][None]))[0]
    gp = np.zeros((world, 2, plan["slots_p"], N), dtype=np.uint64)
    for tt, j in enumerate(own_p):
        mi = sizeQ + j
        gp[rank, 0, tt] = o.ntt_inv(q[mi], roots[mi], e0[len(own_q) + tt])
        gp[rank, 1, tt] = o.ntt_inv(q[mi], roots[mi], e1[len(own_q) + tt])
    all_gather(gp)
    gpf = gp.reshape(world * 2 * plan["slots_p"], N)
    PHatInv, PHat, PInv = c.PHatInvModp(), c.PHatModq(), c.PInvModq()
    rot = np.zeros((2, plan["slots_q"], N), dtype=np.uint64)
    mq = q[own_q]
    for b, e in enumerate((e0, e1)):
        if not own_q:
            continue
        src = np.stack([gpf[plan["gather_p_row"](j, b)] for j in range(sizeP)])
        conv = o.approx_switch_crt_basis(q[sizeQ:], PHatInv, mq, np.ascontiguousarray(PHat[:, own_q]), src)
        for tt, i in enumerate(own_q):
            cv = o.ntt_fwd(q[i], roots[i], conv[tt])
            one = np.array([q[i]], dtype=np.uint64)
            d = o.dcrt_times_scalars(one, o.dcrt_binop(one, 1, e[tt:tt + 1], cv[None]), PInv[i:i + 1])
            if b == 0:
                d = o.dcrt_binop(one, 0, d, ct[0, i][None])
            rot[b, tt] = o.dcrt_automorphism(k, d)[0]
    # ---- collect on rank 0 and compare with the unsharded oracle
    both = np.concatenate([res, rot]).view(np.int64)
    outs = [torch.zeros(4, plan["slots_q"], N, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(outs, torch.from_numpy(np.ascontiguousarray(both)))
    if rank == 0:
        full_res = np.zeros((2, l, N), dtype=np.uint64)
        full_rot = np.zeros((2, sizeQl, N), dtype=np.uint64)
        for r in range(world):
            rows = [i for i in range(sizeQl) if i % world == r]
            lowr = [i for i in rows if i != l]
            got = outs[r].numpy().view(np.uint64)
            full_res[:, lowr] = got[:2, :len(lowr)]
            full_rot[:, rows] = got[2:, :len(rows)]
        want_res = np.stack([c.rescale_poly(sizeQl, ct[0]), c.rescale_poly(sizeQl, ct[1])])
        want_rot = c.automorphism_ct(sizeQl, k, ct, ea, eb)
        result_q.put(bool((full_res == want_res).all()) and bool((full_rot == want_rot).all()))
    c.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,sizeQl", [(2, 6), (3, 5)])
def test_sharded_rescale_and_rotation_over_gloo(world, sizeQl):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world * 10 + sizeQl
    procs = [ctx.Process(target=_rank_rescale_rotate, args=(r, world, port, sizeQl, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_shard_plan_covers_every_limb_once():
    sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
    import ofhe_b200 as m

    for sizeQ, sizeP, sizeQl, world in [(45, 15, 45, 8), (45, 15, 44, 8), (45, 15, 45, 4), (6, 2, 5, 3), (3, 1, 1, 8)]:
        seen_q, seen_p, rows_q, rows_p = [], [], set(), set()
        for r in range(world):
            p = m.shard_plan(sizeQ, sizeP, sizeQl, world, r)
            seen_q += p["own_q"]
            seen_p += p["own_p"]
            assert len(p["own_q"]) <= p["slots_q"] and len(p["own_p"]) <= p["slots_p"]
            for t, i in enumerate(p["own_q"]):
                assert p["gather_q_row"](i) == r * p["slots_q"] + t
                rows_q.add(p["gather_q_row"](i))
            for t, j in enumerate(p["own_p"]):
                assert p["gather_p_row"](j, 1) == (r * 2 + 1) * p["slots_p"] + t
                rows_p.add(p["gather_p_row"](j, 0))
            assert p["own_q"] == p["own_q_full"][:len(p["own_q"])]  # key shards never move between levels
        assert sorted(seen_q) == list(range(sizeQl)) and sorted(seen_p) == list(range(sizeP))
        assert len(rows_q) == sizeQl and len(rows_p) == sizeP
