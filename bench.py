#!/usr/bin/env python
"""bench.py -- CKKS EvalMult + relinearisation throughput (BASELINE.json metric) on N B200s, next to the reference CPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one homomorphic multiplication (EvalMultCore tensor + HYBRID key switch of c2 + the two adds) of two
synthetic full-level ciphertexts at BASELINE.json configs[1]: N = 2^16, depth 44 (45 Q + 15 P limbs), scaleModSize 59,
firstModSize 60, HYBRID with dnum 3.  Prints ONE JSON line (rank 0).

  value        device-resident throughput (inputs already in HBM), mults/s summed over all ranks, CUDA-event timed,
               max over ranks
  e2e          the same operation through the host-buffer C-ABI call (pinned host memory -> device -> host)
  roofline     the dominant kernel family (the limb NTT: strided + tile kernel pair) timed alone on a slab larger
               than L2: algorithmic bytes 16*N per limb-NTT / CUDA-event time vs the measured HBM peak
  cpu_baseline the unmodified reference (oracle/_ref) -- or the C oracle port if that build is absent -- timed on
               this box's host cores on a bounded sample of the same workload
Multi-GPU: ONE multiplication is sharded by RNS limb over the ranks (north_star / SURVEY 8 e) with two NCCL
all-gathers per step; total work per step is fixed -> "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# several independent multiplications (and, with N > 1, their spinning device barriers) are in flight on separate
# streams: give every stream its own hardware queue so that none is serialised behind another one's barrier kernel
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "openfhe-development_b200"))

CFG = dict(N=1 << 16, depth=44, scale_bits=59, first_bits=60, dnum=0, scal_tech=0)
WORKLOAD = "CKKS EvalMult+relinearize, N=2^16, depth 44 (45 Q + 15 P limbs), scaleMod 59, firstMod 60, HYBRID dnum 3"
METRIC = "ckks_evalmult_relin_per_s"
UNIT = "mults/s"
# kernels launched by one EvalMult (hks.cu): tensor 1, INTT 2, base conversion 3, NTT 2, inner product 1,
# ModDown: INTT 2 + base conversion 1 + NTT 2 + (x-conv)*PInv+addend 2
LAUNCHES_PER_STEP = 14  # tensor, INTT x2, bconv (all digits), NTT x2, inner product, INTT x2, bconv, NTT x2, ModDown tail x2


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


def host_topology():
    """(logical CPUs this process may use, physical cores among them)."""
    cpus = sorted(os.sched_getaffinity(0))
    cores = set()
    for c in cpus:
        try:
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/physical_package_id") as f:
                pkg = f.read().strip()
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/core_id") as f:
                cores.add((pkg, f.read().strip()))
        except OSError:
            cores.add(("?", str(c)))
    return len(cpus), max(1, len(cores))


def cpu_reference_child(threads, steps, warmup):
    """Runs in a FRESH process whose OpenMP environment was set by the parent before any libgomp was loaded
    (the reference caches omp_get_max_threads() at static-init time, utils/parallel.h:51-53, and clamps every region
    to it, :112-118): times cc->EvalMult of the unmodified reference (oracle/_ref), or the C oracle port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle, Reference, reference_available, seeded_residues

    N = CFG["N"]
    if reference_available():
        ref = Reference()
        ref.L.ref_set_threads(threads)
        used = int(ref.L.ref_get_threads())
        c = ref.ckks(N, CFG["depth"], CFG["scale_bits"], CFG["first_bits"], CFG["dnum"], CFG["scal_tech"])
        kind = "reference"
    else:
        used = threads
        c = Oracle().ckks(N, CFG["depth"], CFG["scale_bits"], CFG["first_bits"], CFG["dnum"], CFG["scal_tech"])
        kind = "port"
    sizeQl = c.sizeQ
    ct1 = np.stack([seeded_residues(1, c.moduli[:sizeQl], N), seeded_residues(2, c.moduli[:sizeQl], N)])
    ct2 = np.stack([seeded_residues(3, c.moduli[:sizeQl], N), seeded_residues(4, c.moduli[:sizeQl], N)])
    if kind == "reference":
        sec = c.time_evalmult(sizeQl, ct1, ct2, warmup, steps)
    else:
        ea = seeded_residues(5, np.tile(c.moduli, c.dnum), N).reshape(c.dnum, c.sizeQP, N)
        eb = seeded_residues(6, np.tile(c.moduli, c.dnum), N).reshape(c.dnum, c.sizeQP, N)
        for _ in range(warmup):
            c.evalmult(sizeQl, ct1, ct2, ea, eb)
        t0 = time.perf_counter()
        for _ in range(steps):
            c.evalmult(sizeQl, ct1, ct2, ea, eb)
        sec = (time.perf_counter() - t0) / steps
    c.close()
    print(json.dumps(dict(threads=used, sec_per_op=sec, kind=kind)))
    return 0


def cpu_reference_run(steps, warmup, budget_s=150.0):
    """CPU arm: the reference's own EvalMult on this box's host cores, swept over thread counts
    {1, ..., physical cores, all hardware threads} (SURVEY 8 d), every point in a fresh child process with
    OMP_NUM_THREADS / OMP_PROC_BIND / OMP_PLACES set BEFORE the reference library (and its libgomp) is loaded --
    torchrun's exported OMP_NUM_THREADS=1 is overridden.  value = the BEST point; the whole sweep is reported."""
    logical, physical = host_topology()
    cand = sorted({t for t in (1, 8, 16, 32, 64, physical, logical) if 1 <= t <= logical})
    # most informative first, so a tight time budget still sees them: physical cores, all threads, then the rest
    order = [physical, logical] + [t for t in cand if t not in (physical, logical)]
    sweep, kind, t_start = {}, None, time.perf_counter()
    for t in order:
        if sweep and time.perf_counter() - t_start > budget_s:
            break
        env = dict(os.environ)
        env.update(OMP_NUM_THREADS=str(t), OMP_PROC_BIND="close" if t <= physical else "false",
                   OMP_PLACES="cores" if t <= physical else "threads", OMP_DYNAMIC="false")
        env.pop("OMP_THREAD_LIMIT", None)
        n_steps = steps if t > 1 else max(1, min(steps, 2))
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--ref-child", str(t), "--steps", str(n_steps),
                                  "--warmup", str(warmup)], env=env, capture_output=True, text=True, timeout=600)
            rec = json.loads(out.stdout.strip().splitlines()[-1])
        except Exception as ex:  # one failed point does not void the arm
            print(f"[bench] reference child with {t} threads failed: {ex}", file=sys.stderr)
            continue
        sweep[int(rec["threads"])] = 1.0 / rec["sec_per_op"]
        kind = rec["kind"]
    if not sweep:
        raise RuntimeError("no CPU reference measurement succeeded")
    best_t = max(sweep, key=sweep.get)
    sec = 1.0 / sweep[best_t]
    what = "cc->EvalMult of the unmodified reference (oracle/_ref)" if kind == "reference" else "C oracle port"
    return dict(value=sweep[best_t], unit=UNIT, cores=best_t, kind=kind,
                sample=f"{steps} EvalMult calls after {warmup} warm-up on the full workload, {what}; fresh process per "
                       f"thread count with OMP_NUM_THREADS/OMP_PROC_BIND/OMP_PLACES set before load; best of the sweep",
                sweep_mults_per_s={str(k): round(v, 4) for k, v in sorted(sweep.items())},
                host=dict(logical_cpus=logical, physical_cores=physical)), sec


def live_reference_ckks():
    """The unmodified reference's CKKS context for the bench workload (None if oracle/_ref did not travel)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Reference, reference_available

    if not reference_available():
        return None
    ref = Reference()
    return ref.ckks(CFG["N"], CFG["depth"], CFG["scale_bits"], CFG["first_bits"], CFG["dnum"], CFG["scal_tech"])


def committed_profile_numbers():
    """Figures that come from profiler captures, read from the tracked files they were summarised into (never literals
    in this file): DRAM bytes of one roofline-slab forward NTT launch (ncu --set full) and the measured issue cost of
    one warp-butterfly (tools/ubench/bfly.cu)."""
    out = dict(traffic=None, traffic_source=None, cycles_per_warp_butterfly=None, cycles_source=None)
    try:
        with open(os.path.join(ROOT, "profiles", "ntt_roofline_traffic.json")) as f:
            t = json.load(f)
        out["traffic"], out["traffic_source"] = float(t["dram_bytes_per_launch"]), t["source"]
        out["traffic_limbs"], out["traffic_N"] = int(t["limbs"]), int(t["N"])
    except Exception:
        pass
    try:
        with open(os.path.join(ROOT, "profiles", "r02a_ubench_bfly.txt")) as f:
            for ln in f:
                if ln.startswith("V2 fwd"):
                    out["cycles_per_warp_butterfly"] = float(ln.split("=")[1].split()[0])
                    out["cycles_source"] = "profiles/r02a_ubench_bfly.txt (V2: shipped forward butterfly, registers only)"
    except Exception:
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--ref-child", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the config 1 / 3 / 5 side measurements")
    ap.add_argument("--slots", type=int, default=4,
                    help="independent multiplications in flight (streams); the same policy at every GPU count")
    ap.add_argument("--exchange", default="push", choices=["push", "nccl"],
                    help="N>1: limb exchange by NVLink peer stores fused into the INTT (default) or NCCL all-gathers")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="N>1 push mode: eager launches, no CUDA graphs")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if a.ref_child:
        return cpu_reference_child(a.ref_child, max(1, a.steps), max(1, a.warmup))
    if a.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(a.steps, 10))
        warm = max(1, min(a.warmup, 2))
        base, sec = cpu_reference_run(steps, warm)
        line = dict(impl="reference", metric=METRIC, value=base["value"], unit=UNIT, n_gpus=a.gpus, steps=steps,
                    warmup=warm, ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="u64", data="synthetic", config={"workload": WORKLOAD},
                    cpu_baseline=base, e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0,
                                                d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line))
        return 0

    import torch

    import ofhe_b200 as m

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    p = m.ckks_paramgen(CFG["N"], CFG["depth"], CFG["scale_bits"], CFG["first_bits"], CFG["dnum"], CFG["scal_tech"])
    N, sizeQ, sizeP, dnum = p["N"], p["sizeQ"], p["sizeP"], p["dnum"]
    sizeQP = sizeQ + sizeP
    ctx = m.Context(local_rank)
    ring = m.Ring(ctx, N, p["moduli"], p["roots"])
    sizeQl = sizeQ
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_slots = max(1, a.slots)
    streams = [torch.cuda.Stream() for _ in range(n_slots)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def timed(run, steps, warm):
        """W warm-up steps, then exactly `steps` steps between a barrier + synchronize on both sides, CUDA events on
        the launch stream, max over ranks -> milliseconds."""
        run(warm)
        barrier()
        e0.record()
        run(steps)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def fan(fn):
        """run(n): n independent steps issued round-robin over the slot streams (fn(k, slot))."""
        def run(nsteps):
            cur = torch.cuda.current_stream()
            for st in streams:
                st.wait_stream(cur)
            for k in range(nsteps):
                with torch.cuda.stream(streams[k % n_slots]):
                    fn(k, k % n_slots)
            for st in streams:
                cur.wait_stream(st)
        return run

    sampler = ClockSampler(local_rank)
    hks = m.Hks(ring, sizeQ, sizeP, dnum)
    idx2 = list(range(sizeQl)) * 2
    ct1, ct2 = ctx.zeros(2, sizeQl, N), ctx.zeros(2, sizeQl, N)
    ring.fill_uniform(ct1.view(2 * sizeQl, N), idx2, 11)
    ring.fill_uniform(ct2.view(2 * sizeQl, N), idx2, 12)
    # ---- evaluation key: the reference's own relinearisation key when the live reference travelled (single GPU), so
    #      that the timed operation can be checked residue for residue against cc->EvalMult; a seeded key otherwise
    rc = live_reference_ckks() if (world == 1 and rank == 0) else None
    if rc is not None:
        assert [int(v) for v in rc.moduli] == [int(v) for v in p["moduli"]], "parameter generation differs from the reference"
        ka, kb = rc.evalmult_key()
        evk = m.EvalKey(hks, ka, kb)
        del ka, kb
    else:
        evk = m.EvalKey(hks, seed=1234)  # identical on every rank (device PRNG keyed by (seed, row, column))
    parity = None
    closers = []
    extra_fields = {}

    if world == 1:
        # ---- device-resident throughput: independent multiplications round-robin on `slots` streams of ONE context /
        #      ring / key-switch table set / key (per-stream workspace arenas, ofhe_b200.h)
        outs = [ctx.zeros(2, sizeQl, N) for _ in range(n_slots)]
        for st in streams:
            with torch.cuda.stream(st):
                hks.reserve(sizeQl)
        run = fan(lambda k, s_: hks.evalmult(sizeQl, ct1, ct2, evk, outs[s_]))
        run(n_slots)
        torch.cuda.synchronize()
        for s_ in range(1, n_slots):
            assert bool((outs[s_] == outs[0]).all()), "concurrent streams of one context disagree"
        # ---- parity gate BEFORE any timing is accepted (SURVEY 8 d): the same inputs and key through the reference
        if rc is not None:
            want = rc.evalmult(sizeQl, ctx.download(ct1), ctx.download(ct2))
            assert bool((ctx.download(outs[0]) == want).all()), "device EvalMult differs from the reference's cc->EvalMult"
            parity = "bit-exact vs cc->EvalMult of the live unmodified reference (same inputs, same key, full size)"
            rc.close()
        else:
            parity = "reference build absent on this box: cross-stream agreement only"
        sampler.start()
        ms = timed(run, a.steps, max(a.warmup, n_slots))
        sampler.stop_flag.set()
        value = a.steps / (ms * 1e-3)
        parallelism, scaling, launches = (f"single GPU, {n_slots} independent multiplications in flight on {n_slots} "
                                          f"streams of one context"), "weak", LAUNCHES_PER_STEP
        # ---- end to end through the host-buffer entry point (pinned host memory, copies inside the timed region)
        h1 = torch.empty(2, sizeQl, N, dtype=torch.int64).pin_memory()
        h2 = torch.empty(2, sizeQl, N, dtype=torch.int64).pin_memory()
        ho = torch.empty(2, sizeQl, N, dtype=torch.int64).pin_memory()
        h1.copy_(ct1.cpu())
        h2.copy_(ct2.cpu())
        e2e_steps = max(6, min(a.steps, 32))
        hos = [torch.empty(2, sizeQl, N, dtype=torch.int64).pin_memory() for _ in range(min(e2e_steps, 6))]
        hos = [hos[i % len(hos)] for i in range(e2e_steps)]  # results cycle through 6 pinned buffers
        hks.evalmult_host(sizeQl, h1, h2, evk, ho)
        hks.evalmult_host_batch(sizeQl, [h1] * 3, [h2] * 3, evk, hos[:3])
        barrier()
        t0 = time.perf_counter()
        hks.evalmult_host_batch(sizeQl, [h1] * e2e_steps, [h2] * e2e_steps, evk, hos)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        assert bool((ho.cuda() == outs[0]).all()), "host-buffer path and device path disagree"
        for hb in (hos[0], hos[-1]):
            assert bool((hb == ho).all()), "pipelined host path disagrees with the single-call host path"
    else:
        # ---- limb-sharded across the ranks (SURVEY 8 e): ONE multiplication is split by RNS limb.  `slots` pipeline
        #      slots on `slots` streams (independent ciphertext pairs) let one multiplication's exchange overlap
        #      another's kernels -- the same in-flight policy as the single-GPU run.
        shards = [m.Shard(ring, sizeQ, sizeP, dnum, rank, world) for _ in range(n_slots)]
        info = shards[0].info(sizeQl)
        own_q, own_p = shards[0].owned(sizeQl)
        nq = len(own_q)
        key_idx = [i for i in range(sizeQ) if i % world == rank] + [sizeQ + j for j in own_p]
        # this rank's rows of the (identical everywhere) key and inputs
        sel = torch.tensor(key_idx, device="cuda")
        ea = evk_rows(m, ctx, hks, evk, dnum, sizeQP, N, 0).index_select(1, sel).contiguous()
        eb = evk_rows(m, ctx, hks, evk, dnum, sizeQP, N, 1).index_select(1, sel).contiguous()
        selq = torch.tensor(own_q, device="cuda")
        s1, s2 = ct1.index_select(1, selq).contiguous(), ct2.index_select(1, selq).contiguous()
        outs = [ctx.zeros(2, nq, N) for _ in range(n_slots)]

        def exchange_obj(obj):
            objs = [None] * world
            dist.all_gather_object(objs, obj)
            return objs

        graphs = [None] * n_slots
        if a.exchange == "push":
            pqs = [m.PeerBuffers(ctx, world * info["slots_q"] * N, rank, world, exchange_obj) for _ in range(n_slots)]
            pps = [m.PeerBuffers(ctx, world * 2 * info["slots_p"] * N, rank, world, exchange_obj) for _ in range(n_slots)]
            pfl = m.PeerBuffers(ctx, m.lib().ofhe_peer_flags_words(world, 2 * n_slots), rank, world, exchange_obj)
            closers += pqs + pps + [pfl]
            dist.barrier()

            def one(s, c1=None, c2=None):
                shards[s].evalmult_push(sizeQl, s1 if c1 is None else c1, s2 if c2 is None else c2, ea, eb, outs[s],
                                        pqs[s].array, pps[s].array, pfl.array, slot=s)
        else:
            gqs = [ctx.zeros(world, info["slots_q"], N) for _ in range(n_slots)]
            gps = [ctx.zeros(world, 2, info["slots_p"], N) for _ in range(n_slots)]

            def all_gather(buf):
                dist.all_gather_into_tensor(buf, buf[rank])  # in place: the send segment is this rank's slice of buf

            def one(s, c1=None, c2=None):
                shards[s].evalmult(sizeQl, s1 if c1 is None else c1, s2 if c2 is None else c2, ea, eb, outs[s],
                                   gqs[s], gps[s], all_gather)

        def step(k, s):
            if graphs[s] is not None:
                graphs[s].replay()
            else:
                one(s)

        run = fan(step)
        run(max(a.warmup, n_slots))
        if a.exchange == "push" and a.graph:
            # one CUDA graph per pipeline slot (all launches of a multiplication incl. the device barriers, whose epochs
            # live in device memory): removes the per-launch host cost that dominates once the work is split 8 ways
            torch.cuda.synchronize()
            dist.barrier()
            try:
                for s_ in range(n_slots):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=streams[s_]):
                        one(s_)
                    graphs[s_] = g
            except Exception as ex:  # capture unsupported: stay on eager launches
                graphs = [None] * n_slots
                print(f"[bench] CUDA graph capture failed, eager launches: {ex}", file=sys.stderr)
            torch.cuda.synchronize()
            dist.barrier()
        # ---- parity gate: every rank's owned rows -> rank 0, compared with the UNSHARDED device EvalMult of the same
        #      inputs and key (itself pinned to the reference by the single-GPU run and tests/)
        run(n_slots)
        torch.cuda.synchronize()
        pad = ctx.zeros(2, info["slots_q"], N)
        ok = 1
        for s_ in range(n_slots):
            pad[:, :nq] = outs[s_]
            allp = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, allp, dst=0)
            if rank == 0:
                if s_ == 0:
                    want = ctx.zeros(2, sizeQl, N)
                    hks.evalmult(sizeQl, ct1, ct2, evk, want)
                    torch.cuda.synchronize()
                for r_ in range(world):
                    rows = [i for i in range(sizeQl) if i % world == r_]
                    if not bool((allp[r_][:, :len(rows)] == want[:, rows]).all()):
                        ok = 0
        flag = torch.tensor([ok], device="cuda")
        dist.broadcast(flag, src=0)
        assert int(flag.item()) == 1, "limb-sharded EvalMult differs from the unsharded device EvalMult"
        parity = (f"every rank's rows of all {n_slots} pipeline slots gathered to rank 0: bit-exact vs the unsharded "
                  f"device EvalMult (same inputs, same key)")
        sampler.start()
        ms = timed(run, a.steps, max(a.warmup, n_slots))
        sampler.stop_flag.set()
        value = a.steps / (ms * 1e-3)
        how = ("exchange by NVLink peer stores fused into the INTT + device barriers"
               + (", one CUDA graph per slot" if graphs[0] is not None else "")) if a.exchange == "push" \
            else "2 NCCL all-gathers per EvalMult"
        parallelism, scaling = f"RNS limbs sharded over {world} GPUs, {how}, {n_slots} pipeline slots", "strong"
        launches = 18 if a.exchange == "push" else LAUNCHES_PER_STEP
        if a.exchange == "push":
            torch.cuda.synchronize()
            dist.barrier()
            assert ctx.peer_barrier_timeouts(pfl.local, world, 2 * n_slots) == 0, "device barrier timed out"
        # ---- the other mode SURVEY 8(e) asks to report: REPLICAS -- every GPU multiplies its own ciphertext pairs
        #      with a full key copy, no communication (weak scaling of the same operation)
        routs = [ctx.zeros(2, sizeQl, N) for _ in range(n_slots)]
        for st in streams:
            with torch.cuda.stream(st):
                hks.reserve(sizeQl)
        rrun = fan(lambda k, s_: hks.evalmult(sizeQl, ct1, ct2, evk, routs[s_]))
        rms = timed(rrun, a.steps, max(a.warmup, n_slots))
        extra_fields["replicas"] = dict(value=world * a.steps / (rms * 1e-3), unit=UNIT, scaling="weak",
                                        note=f"{world} GPUs x independent unsharded EvalMults ({a.steps} per GPU), "
                                             f"max-over-ranks device time {rms:.3f} ms")
        # ---- end to end: every rank moves its own limb rows host<->device around the sharded multiplication
        h1 = torch.empty(2, nq, N, dtype=torch.int64).pin_memory()
        h2 = torch.empty(2, nq, N, dtype=torch.int64).pin_memory()
        ho = torch.empty(2, nq, N, dtype=torch.int64).pin_memory()
        h1.copy_(s1.cpu())
        h2.copy_(s2.cpu())
        d1, d2 = torch.empty_like(s1), torch.empty_like(s2)
        e2e_steps = max(3, min(a.steps, 10))

        def e2e_step():
            d1.copy_(h1, non_blocking=True)
            d2.copy_(h2, non_blocking=True)
            one(0, d1, d2)
            ho.copy_(outs[0], non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        closers = shards + closers
    ct_bytes = 2 * sizeQl * N * 8

    # ---- roofline of the dominant kernel family: forward limb-NTT on a slab larger than L2 ---------------------
    L = 240  # 240 limbs x 512 KiB = 120 MiB in place, > L2; 4 limbs per modulus in modulus-major order (as the key
    slab = ctx.zeros(L, N)  # switch orders its multi-polynomial launches)
    lidx = [i // (L // sizeQP) for i in range(L)]
    ring.fill_uniform(slab, lidx, 99)
    for _ in range(3):
        ring.ntt_fwd(slab, slab, lidx)
    torch.cuda.synchronize()
    reps = 10
    e0.record()
    for _ in range(reps):
        ring.ntt_fwd(slab, slab, lidx)
    e1.record()
    torch.cuda.synchronize()
    ntt_ms = e0.elapsed_time(e1) / reps
    peak, peak_src = measured_peak_gbs()
    ntt_bytes = 16.0 * N * L
    achieved = ntt_bytes / (ntt_ms * 1e-3) / 1e9
    prof = committed_profile_numbers()
    traffic = prof["traffic"] if (prof["traffic"] and prof.get("traffic_limbs") == L and prof.get("traffic_N") == N) else None
    roofline = dict(bound="hbm", kernel="forward limb-NTT (ntt_fwd_strided<4> + ntt_fwd_tile<4>)", achieved=achieved,
                    peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, traffic_source=prof["traffic_source"],
                    peak_source=peak_src, algorithmic_bytes_per_launch=ntt_bytes, us_per_limb_ntt=ntt_ms * 1e3 / L)
    # the co-roof that binds a 60-bit NTT on sm_100: integer issue, measured per warp-butterfly by
    # tools/ubench/bfly.cu on registers only
    if prof["cycles_per_warp_butterfly"]:
        sm_mhz = (sampler.summary().get("sm_max_mhz") or 1965.0)
        logn = N.bit_length() - 1
        cyc = prof["cycles_per_warp_butterfly"]
        floor_us = (N // 2 * logn / 32.0) * cyc / (ctx.sm_count * 4) / sm_mhz
        roofline["int_issue"] = dict(bound="integer issue (IMAD.WIDE/IMAD + IADD3 pipes), registers-only microbenchmark",
                                     cycles_per_warp_butterfly=cyc, source=prof["cycles_source"],
                                     floor_us_per_limb_ntt=floor_us, frac=floor_us / (ntt_ms * 1e3 / L))
    del slab

    line = None
    aux = None
    if rank == 0 and not a.no_aux and world == 1:
        try:
            aux = aux_measurements(m, torch, ctx)
        except Exception as ex:  # side measurements never void the headline
            aux = {"error": repr(ex)}
    if world > 1 and not a.no_aux:  # BASELINE config 5 at N GPUs (every rank takes part)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import aux_bench

            aux = {"config5_ntt_sweep_limb_sharded": aux_bench.config5_sharded(m, torch, ctx, dist, rank, world)}
        except Exception as ex:
            aux = {"error": repr(ex)}
    if rank == 0:
        base = None
        if not a.no_cpu_baseline and world == 1:  # the CPU arm is timed at N=1 only (the ranks would fight for the cores)
            base, _ = cpu_reference_run(a.cpu_steps, 1, budget_s=90.0)
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=a.warmup,
                    ms_per_step=ms / a.steps, higher_is_better=True, scaling=scaling, vs_baseline=None, dtype="u64",
                    data="synthetic",
                    config={"workload": WORKLOAD, "parallelism": parallelism,
                            "l2": "per-step working set ~0.6 GB (evk 180 MiB + digits 90 MiB + ...) exceeds the 126 MB L2"},
                    parity_checked=parity,
                    e2e=dict(value=e2e_steps / e2e_s, unit=UNIT, h2d_bytes_per_step=2 * ct_bytes,
                             d2h_bytes_per_step=ct_bytes, steps=e2e_steps),
                    gpu_launches=launches * a.steps, roofline=roofline, cpu_baseline=base,
                    clocks=sampler.summary())
        line.update(extra_fields)
        if aux is not None:
            line["aux"] = aux
        print(json.dumps(line))
    for o in closers + [evk, hks, ring, ctx]:
        o.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def evk_rows(m, ctx, hks, evk, dnum, sizeQP, N, which):
    """Device copy [dnum][sizeQP][N] of the key's A (which=0) or B (which=1) vector."""
    t = ctx.zeros(dnum, sizeQP, N)
    evk.copy_to(which, t)
    return t


def aux_measurements(m, torch, ctx):
    """BASELINE.json configs 1, 3 and 5 beside the headline (single GPU): see tools/aux_bench.py."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import aux_bench

    return aux_bench.run(m, torch, ctx)


if __name__ == "__main__":
    sys.exit(main())
