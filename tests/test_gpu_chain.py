"""Device-resident ciphertext lifecycle: a chain EvalMult -> Rescale -> EvalRotate -> EvalAdd ->
EvalMult -> Rescale whose intermediates never leave the GPU equals the oracle's chain bit for bit, and the whole chain
replays from ONE CUDA graph (all launches are capturable: no allocation, copy to host or synchronisation inside)."""
import numpy as np
import pytest

from oracle_lib import seeded_residues

pytestmark = pytest.mark.gpu


def test_chain_device_resident_and_graph_replay(gpu, oracle):
    import torch

    ctx, m = gpu
    args = (1 << 12, 6, 50, 60, 3, 0)
    c = oracle.ckks(*args)
    try:
        p = m.ckks_paramgen(*args)
        N, L = p["N"], p["sizeQ"]
        ring = m.Ring(ctx, N, p["moduli"], p["roots"])
        hks = m.Hks(ring, L, p["sizeP"], p["dnum"])
        ea = seeded_residues(61, np.tile(c.moduli, c.dnum), N).reshape(c.dnum, c.sizeQP, N)
        eb = seeded_residues(62, np.tile(c.moduli, c.dnum), N).reshape(c.dnum, c.sizeQP, N)
        evk = m.EvalKey(hks, ea, eb)
        ct1 = np.stack([seeded_residues(63, c.moduli[:L], N), seeded_residues(64, c.moduli[:L], N)])
        ct2 = np.stack([seeded_residues(65, c.moduli[:L], N), seeded_residues(66, c.moduli[:L], N)])
        k = m.find_automorphism_index(2, 2 * N)
        # ---- oracle chain
        q = c.moduli
        m1 = c.evalmult(L, ct1, ct2, ea, eb)
        r1 = np.stack([c.rescale_poly(L, m1[0]), c.rescale_poly(L, m1[1])])
        rot = c.automorphism_ct(L - 1, k, r1, ea, eb)
        s = np.stack([oracle.dcrt_binop(q[:L - 1], 0, rot[0], r1[0]), oracle.dcrt_binop(q[:L - 1], 0, rot[1], r1[1])])
        m2 = c.evalmult(L - 1, s, r1, ea, eb)
        want = np.stack([c.rescale_poly(L - 1, m2[0]), c.rescale_poly(L - 1, m2[1])])
        # ---- device chain: every intermediate is a device slab
        d1, d2 = ctx.upload(ct1), ctx.upload(ct2)
        b_m1, b_r1 = ctx.zeros(2, L, N), ctx.zeros(2, L - 1, N)
        b_rot, b_s = ctx.zeros(2, L - 1, N), ctx.zeros(2, L - 1, N)
        b_m2, out = ctx.zeros(2, L - 1, N), ctx.zeros(2, L - 2, N)
        idx2 = list(range(L - 1)) * 2

        def chain():
            hks.evalmult(L, d1, d2, evk, b_m1)
            hks.rescale(L, b_m1, b_r1, 2)
            hks.rotate(L - 1, k, b_r1, evk, b_rot)
            ring.binop(0, b_s.view(2 * (L - 1), N), b_rot.view(2 * (L - 1), N), b_r1.view(2 * (L - 1), N), idx2)
            hks.evalmult(L - 1, b_s, b_r1, evk, b_m2)
            hks.rescale(L - 1, b_m2, out, 2)

        chain()
        torch.cuda.synchronize()
        assert (ctx.download(out) == want).all()
        # ---- the same chain as one CUDA graph
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            chain()  # warm-up on the capture stream (tables, scratch and the automorphism map already exist)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            chain()
        for _ in range(2):
            out.zero_()
            b_m1.zero_()
            g.replay()
            torch.cuda.synchronize()
            assert (ctx.download(out) == want).all()
        for o in (evk, hks, ring):
            o.close()
    finally:
        c.close()
