"""Real multi-GPU check of the limb-sharded EvalMult over NCCL (launch with torchrun, one rank per GPU):
every rank runs its share, rank 0 gathers the owned rows and compares with the unsharded device result.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ofhe_b200 as m  # noqa: E402
from oracle_lib import seeded_residues  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
logn = int(os.environ.get("LOGN", "13"))
depth = int(os.environ.get("DEPTH", "8"))
p = m.ckks_paramgen(1 << logn, depth, 59, 60, 0, 0)
N, sizeQ, sizeP, dnum = p["N"], p["sizeQ"], p["sizeP"], p["dnum"]
ctx = m.Context(local)
ring = m.Ring(ctx, N, p["moduli"], p["roots"])
ok = True
for sizeQl in (sizeQ, sizeQ - 1):
    ea = seeded_residues(81, np.tile(p["moduli"], dnum), N).reshape(dnum, sizeQ + sizeP, N)
    eb = seeded_residues(82, np.tile(p["moduli"], dnum), N).reshape(dnum, sizeQ + sizeP, N)
    ct1 = np.stack([seeded_residues(83, p["moduli"][:sizeQl], N), seeded_residues(84, p["moduli"][:sizeQl], N)])
    ct2 = np.stack([seeded_residues(85, p["moduli"][:sizeQl], N), seeded_residues(86, p["moduli"][:sizeQl], N)])
    sh = m.Shard(ring, sizeQ, sizeP, dnum, rank, world)
    info = sh.info(sizeQl)
    own_q, own_p = sh.owned(sizeQl)
    plan = m.shard_plan(sizeQ, sizeP, sizeQl, world, rank)
    key_rows = plan["own_q_full"] + [sizeQ + j for j in own_p]
    d1, d2 = ctx.upload(ct1[:, own_q]), ctx.upload(ct2[:, own_q])
    out = ctx.zeros(2, len(own_q), N)
    a_d, b_d = ctx.upload(ea[:, key_rows]), ctx.upload(eb[:, key_rows])
    gq, gp = ctx.zeros(world, info["slots_q"], N), ctx.zeros(world, 2, info["slots_p"], N)

    def all_gather(buf):
        dist.all_gather_into_tensor(buf, buf[rank].clone() if False else buf[rank])  # in place: NCCL allows send == recv segment

    sh.evalmult(sizeQl, d1, d2, a_d, b_d, out, gq, gp, all_gather)
    torch.cuda.synchronize()
    # the same multiplication with the exchange done by peer stores + device barriers (no collective), twice in a row
    # on the same buffers (buffer-reuse safety), must reproduce the NCCL result bit for bit
    def exchange(obj):
        objs = [None] * world
        dist.all_gather_object(objs, obj)
        return objs

    pq = m.PeerBuffers(ctx, world * info["slots_q"] * N, rank, world, exchange)
    pp = m.PeerBuffers(ctx, world * 2 * info["slots_p"] * N, rank, world, exchange)
    pf = m.PeerBuffers(ctx, m.lib().ofhe_peer_flags_words(world, 2), rank, world, exchange)
    dist.barrier()
    out2 = ctx.zeros(2, len(own_q), N)
    for _ in range(2):
        sh.evalmult_push(sizeQl, d1, d2, a_d, b_d, out2, pq.array, pp.array, pf.array, slot=0)
    torch.cuda.synchronize()
    dist.barrier()
    timeouts = ctx.peer_barrier_timeouts(pf.local, world, 2)
    push_same = bool((out2 == out).all()) and timeouts == 0
    print(f"shard_check push world={world} rank={rank} sizeQl={sizeQl}: {'OK' if push_same else 'MISMATCH'}"
          f" (barrier timeouts {timeouts})", flush=True)
    ok = ok and push_same
    for b in (pq, pp, pf):
        b.close()
    # assemble on rank 0: pad every rank's rows to slots_q
    pad = ctx.zeros(2, info["slots_q"], N)
    pad[:, :len(own_q)] = out
    allp = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(allp, pad)
    if rank == 0:
        full = np.zeros((2, sizeQl, N), dtype=np.uint64)
        for r in range(world):
            rows = [i for i in range(sizeQl) if i % world == r]
            full[:, rows] = ctx.download(allp[r])[:, :len(rows)]
        hks = m.Hks(ring, sizeQ, sizeP, dnum)
        evk = m.EvalKey(hks, ea, eb)
        ref = ctx.zeros(2, sizeQl, N)
        hks.evalmult(sizeQl, ctx.upload(ct1), ctx.upload(ct2), evk, ref)
        same = bool((ctx.download(ref) == full).all())
        print(f"shard_check world={world} N=2^{logn} sizeQl={sizeQl}: {'OK' if same else 'MISMATCH'}", flush=True)
        ok = ok and same
        evk.close()
        hks.close()
    sh.close()
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
