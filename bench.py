#!/usr/bin/env python
"""bench.py -- short-range force throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one whole short-range force evaluation (device tree build + P2M/M2M + fused list walk and P2P
incl. the 26 periodic images + M2L + L2L/L2P) of a synthetic LambdaCDM-like periodic particle set.
The workload is STRONG-scaled: --npart-side^3 particles in total (default 512^3, the size the metric is
quoted on), split over the N ranks by the reference's domain decomposition.
One JSON line on stdout (rank 0); see DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "photons-2.0_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

OPS_PER_INTERACTION = 24          # FMA-pipe ops of the P2P inner loop
METRIC = "short-range force: particles/s per force step"
_OUT = sys.stdout


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--npart-side", type=int, default=512, help="particles per dimension (total, all GPUs)")
    ap.add_argument("--nside", type=int, default=0, help="PM mesh side NSIDE (0: = npart-side)")
    ap.add_argument("--maxleaf", type=int, default=8)
    ap.add_argument("--theta", type=float, default=0.4)
    ap.add_argument("--disp-rms", type=float, default=0.3, help="rms Zel'dovich displacement in grid spacings")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--ic", default="lcdm", choices=["lcdm", "poisson", "merger"],
                    help="lcdm: grid + Zel'dovich displacements (default); poisson: uniform random (the reference's ic_uniform); "
                         "merger: the reference's demo/ic_merger.gdt2 (non-periodic Newtonian build, M2L-heavy; 1 GPU)")
    ap.add_argument("--cpu-sample-side", type=int, default=96, help="particles per dimension of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--pm", action="store_true", help="also time the PM long-range force (pn2_pm_force_device, NSIDE mesh) on the same particles")
    ap.add_argument("--rebalance", type=int, default=0,
                    help="after the timed region (N > 1): this many iterations of the reference's load-balance loop -- DTIME_FRACTION from the "
                         "ranks' list sizes, determine_split_domtree, particle migration, force step -- reported as `rebalance`")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa(index):
    """Run this process (and first-touch its pinned buffers) on the NUMA node the GPU hangs off: with 8 ranks copying
    at once, remote-node pinned memory is what limits the host-pointer path (e2e).  Best effort; returns a description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(hnd).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = int(open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node").read())
        if node < 0:
            return "numa_node unknown"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return f"node {node}: no allowed cpus"
        os.sched_setaffinity(0, allowed)
        return f"node {node}, {len(allowed)} cpus"
    except Exception as ex:
        return "unbound (" + repr(ex)[:80] + ")"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_cpu_run(args, side, nranks, repeat=1, keep=False):
    """The reference's own CPU implementation of the path (oracle/_ref/ref_fmm: the UNMODIFIED reference
    sources with a fork/socketpair MPI shim, PM excluded) on a bounded sample of the workload: the same
    generator at side^3 particles with NSIDE = side * (nside / npart_side), i.e. the same interactions per
    particle.  Returns (particles/s, interactions/s, seconds per force evaluation, kind)."""
    import torch
    import synthetic
    from oracle import pn_ref, pn_oracle
    nside_pm = max(2, int(round(side * (args.nside or args.npart_side) / args.npart_side)))
    # The reference's LET exchange assumes that a domain meets every peer at most once through the periodic wrap: its run
    # does not end when a domain is thinner than twice the cut-off radius (observed: 32^3 / NSIDE 32 at 12 or 16 ranks,
    # 40^3 at 16).  Halve the rank count until every domain of src/domains.c's decomposition is wide enough.
    import domains
    cutoff = 4.5 * 1.25 * synthetic.BOX / nside_pm                      # src/initial.c:316-345
    while nranks > 1 and min(float(min(d.hi[k] - d.lo[k] for k in range(3))) for d in domains.domain_boxes(nranks, synthetic.BOX)) < 2.1 * cutoff:
        nranks //= 2
    pos = synthetic.lcdm_like(side, disp_rms=args.disp_rms, seed=12345, device="cpu").numpy()
    n = len(pos)
    mass = synthetic.particle_mass(n)
    if pn_ref.available():
        kind = "reference"
        t0 = time.perf_counter()
        ranks = pn_ref.run_reference(pos, synthetic.BOX, nside_pm, mass, maxleaf=args.maxleaf, theta=args.theta, nranks=nranks,
                                     capture=1, repeat=repeat, timeout=3000)   # capture=1: the harness counts local interactions
        wall = time.perf_counter() - t0
        # the harness times construct+prepare+task+ext per rank (max over ranks = the step)
        sec = max(r["timing"]["total"] for r in ranks) / max(1, repeat)
        nint = sum(r["nint_local"] + r["p2p_count_remote"] for r in ranks)
        if not (sec > 0):
            sec = wall / max(1, repeat)
        if keep:           # the sample's positions and the reference's accelerations, for the parity block of the bench line
            kept = {"pos": pos, "acc": pn_ref.gather_acc(ranks, n), "nleaf": [r["last_leaf"] - r["first_leaf"] for r in ranks],
                    "nint": [r["nint_local"] + r["p2p_count_remote"] for r in ranks], "mass": mass}
    else:
        kind = "port"
        prm = pn_oracle.make_params(synthetic.BOX, nside_pm, n, mass, maxleaf=args.maxleaf, theta=args.theta)
        t0 = time.perf_counter()
        _, cnt = pn_oracle.force(pos, prm, 1)
        sec = time.perf_counter() - t0
        nint = cnt["int_local"] + cnt["int_remote"]
        nranks = 1
    out = {"pps": n / sec, "ips": nint / sec, "sec": sec, "kind": kind, "cores": nranks, "n": n, "side": side,
           "nside": nside_pm, "interactions_per_particle": nint / n}
    if keep and kind == "reference":
        out["kept"] = kept
    return out


def parity_vs_reference(args, r):
    """The product path against the UNMODIFIED reference on the CPU-baseline sample of this very run: Mode B in both
    arithmetic modes at the reference's rank count (one context per rank on this GPU, LET blocks exchanged
    device-to-device by pn2_exchange_local: same domain rule, src/domains.c:399-428), accelerations compared with
    the ones the reference harness computed (rms relative error, north_star: <= 1e-6 FP64, <= 1e-4 FP32), leaf
    and interaction counts compared rank by rank."""
    import pn2gpu
    import domains
    import synthetic
    k = r["kept"]
    pos, ref, nr = k["pos"], k["acc"], r["cores"]
    doms = domains.domain_boxes(nr, synthetic.BOX)
    owner = domains.domain_of(pos, nr, synthetic.BOX)
    idx = [np.nonzero(owner == q)[0] for q in range(nr)]
    out = {"sample": f"{r['side']}^3 particles, NSIDE {r['nside']}, {nr} ranks (the cpu_baseline run of this line)",
           "reference": "oracle/_ref/ref_fmm (unmodified reference sources)", "tolerance": {"fp64_rms": 1e-6, "fp32_rms": 1e-4}}
    for name, precision in (("fp64", pn2gpu.FP64), ("fp32", pn2gpu.FP32)):
        prm = pn2gpu.make_params(synthetic.BOX, r["nside"], len(pos), k["mass"], maxleaf=args.maxleaf, theta=args.theta, precision=precision)
        ctxs = [pn2gpu.Context(prm) for _ in range(nr)]
        accs = pn2gpu.force_step_local_ranks(ctxs, [pos[i] for i in idx], doms)
        acc = np.zeros_like(pos)
        for q in range(nr):
            acc[idx[q]] = accs[q]
        infos = [c.step_info() for c in ctxs]
        for c in ctxs:
            c.close()
        out[name + "_rms"] = float(np.sqrt(((acc - ref) ** 2).sum() / (ref ** 2).sum()))
        out["nleaf_equal"] = bool(out.get("nleaf_equal", True) and [i["nleaf"] for i in infos] == [int(x) for x in k["nleaf"]])
        out["nint_equal"] = bool(out.get("nint_equal", True) and [i["n_interactions"] for i in infos] == [int(x) for x in k["nint"]])
    out["pass"] = bool(out["fp64_rms"] <= 1e-6 and out["fp32_rms"] <= 1e-4 and out["nleaf_equal"] and out["nint_equal"])
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    nranks = 1
    while nranks * 2 <= cores:
        nranks *= 2                                   # the shim runs any count; powers of two keep domains cubic
    side = args.cpu_sample_side
    for _ in range(args.warmup and 1):
        reference_cpu_run(args, 32, min(nranks, 8))   # warm the page cache / binary, untimed
    t0 = time.perf_counter()
    res = [reference_cpu_run(args, side, nranks) for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    sec = float(np.mean([r["sec"] for r in res]))
    r = res[0]
    pps = r["n"] / sec
    sample = (f"{side}^3 particles of the same generator, NSIDE {r['nside']}, one force evaluation per step at NP={r['cores']} "
              f"ranks (fork/socketpair MPI shim), {r['interactions_per_particle']:.0f} interactions/particle")
    out = {"metric": METRIC, "value": pps, "unit": "particles/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "impl": "reference",
           "config": workload_config(args, args.gpus) | {
               "workload": f"bounded sample of the arm's workload: synthetic LCDM-like {side}^3 particles (same generator, seed), NSIDE {r['nside']}, "
                           f"Zel'dovich displacement rms {args.disp_rms} grid spacings",
               "npart": r["n"], "nside": r["nside"], "parallelism": f"mpi_shim_ranks{r['cores']}", "precision_mode": "fp64 (the reference's arithmetic)",
               "sample_of": f"{args.npart_side}^3 particles, NSIDE {args.nside or args.npart_side} (the GPU arm's workload; same interactions per particle)",
               "cache": "n/a (host)", "reference_sample": sample},
           "p2p_ginteractions_per_s": r["interactions_per_particle"] * pps / 1e9,
           "cpu_baseline": {"value": pps, "unit": "particles/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
           "e2e": {"value": pps, "unit": "particles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": wall}
    print(json.dumps(out), file=_OUT, flush=True)


def workload_config(args, ngpu):
    nside = args.nside or args.npart_side
    if args.ic == "merger":
        return {"workload": "the reference's demo/ic_merger.gdt2 (60000 particles, two clustered haloes) shifted into a BOX 400, non-periodic "
                            "Newtonian build (no -DPERIODIC_CONDITION -DLONGSHORT), NSIDE 8", "npart": 60000, "nside": 8, "maxleaf": args.maxleaf,
                "theta": args.theta, "precision_mode": args.precision, "parallelism": f"domains{ngpu}", "cache": "fits L2"}
    what = (f"synthetic LCDM-like {args.npart_side}^3 particles, periodic box {100000.0:g} kpc/h, NSIDE {nside}, "
            f"Zel'dovich displacement rms {args.disp_rms} grid spacings, seed 12345") if args.ic == "lcdm" else \
           (f"Poisson (uniform random, the reference's ic_uniform: src/initial.c:558-618) {args.npart_side}^3 particles, periodic box "
            f"{100000.0:g} kpc/h, NSIDE {nside}, seed 378412")
    return {"workload": what,
            "npart": args.npart_side ** 3, "nside": nside, "maxleaf": args.maxleaf, "theta": args.theta,
            "precision_mode": args.precision, "parallelism": f"domains{ngpu}",
            "cache": "inputs (positions, tree, multipoles: several GB) exceed the 126 MB L2; no explicit flush"}


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: anything libraries print to fd 1 (e.g. "NCCL version ...") goes to stderr
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference_arm(args)
    import torch
    import torch.distributed as dist
    import pn2gpu
    import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if world > 1 else "not bound (one rank)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nside = args.nside or args.npart_side
    ntot = args.npart_side ** 3
    mass = synthetic.particle_mass(ntot)
    precision = pn2gpu.FP32 if args.precision == "fp32" else pn2gpu.FP64
    box = synthetic.BOX
    if args.ic == "merger":
        if world > 1:
            raise SystemExit("bench.py: --ic merger is a single-GPU line")
        g = np.load(os.path.join(ROOT, "tests", "golden", "merger_open_np1.npz"))
        mpos = np.load(os.path.join(ROOT, "tests", "golden", "merger_pos_f32.npy")).astype(np.float64) + float(g["shift"])
        box, nside, ntot, mass = float(g["box"]), int(g["nside"]), len(mpos), float(g["mass"])
        prm = pn2gpu.make_params(box, nside, ntot, mass, maxleaf=args.maxleaf, theta=args.theta, soft=float(g["soft"]), periodic=0, longshort=0,
                                 precision=precision)
    else:
        prm = pn2gpu.make_params(box, nside, ntot, mass, maxleaf=args.maxleaf, theta=args.theta, precision=precision)
    ctx = pn2gpu.Context(prm, device=local)

    # ---- workload: every rank generates the same field, keeps the particles of its own domain ----
    if args.ic == "merger":
        pos_all = torch.from_numpy(mpos).to(dev)
    elif args.ic == "poisson":
        pos_all = synthetic.poisson(ntot, seed=378412, device=dev)
    else:
        pos_all = synthetic.lcdm_like(args.npart_side, disp_rms=args.disp_rms, seed=12345, device=dev)
    migrate = None
    if world > 1:
        # every rank starts with an arbitrary slice of the set; the particles reach the rank that owns them under the
        # reference's domain tree through the device migration (pn2_migrate_*: owner rule + all-to-all-v over NCCL)
        import domains
        dtree = domains.DomainTree(world, synthetic.BOX)
        doms = dtree.boxes()
        dom = doms[rank]
        ctx.set_comm_torch(rank, world, doms)
        held = pos_all[rank::world].contiguous()
        n_own = int((domains.owner_of(pos_all, world, dtree.splits) == rank).sum())
        del pos_all
        torch.cuda.empty_cache()
        torch.cuda.synchronize()
        dist.barrier()
        for it in range(2):                    # the first pass allocates the pools and opens the NCCL channels
            ctx.sync()
            dist.barrier()
            ctx.timer_start(2)
            ctx.migrate_begin(held.data_ptr(), 3, held.shape[0], dtree.splits)
            ctx.migrate_exchange_nccl()
            mig_ms = ctx.timer_stop(2)
        _, n_new, _ = ctx.migrate_result()
        if n_new != n_own:
            raise SystemExit(f"bench.py: rank {rank} received {n_new} particles, owns {n_own}")
        pos = torch.empty((n_new, 3), dtype=torch.float64, device=dev)
        pn2gpu._ck(pn2gpu.lib().pn2_migrate_fetch(ctx.h, pos.data_ptr()))
        tm = torch.tensor([mig_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        migrate = {"ms": float(tm[0]), "records": ntot, "record_bytes": 24,
                   "gbytes_per_s": ntot * 24 / (float(tm[0]) * 1e-3) / 1e9,
                   "what": "pn2_migrate_begin + pn2_migrate_exchange_nccl of all particles from a round-robin start (setup, not in the timed step)"}
        del held
        pos_all = None
    else:
        pos = pos_all
        dom = pn2gpu.make_domain([0, 0, 0], [box] * 3, 0)
    del pos_all
    torch.cuda.empty_cache()
    n = pos.shape[0]
    acc = torch.empty_like(pos)
    torch.cuda.synchronize()

    def step():
        ctx.force_step_device(pos.data_ptr(), n, acc.data_ptr(), dom)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    fma_peak = ctx.fma_peak(precision != pn2gpu.FP32)     # FFMA issue rate (FP32 mode) / DFMA issue rate (FP64 mode)
    dfma_peak = ctx.fma_peak(True)                        # the multipole operators always run in FP64
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    phase = []
    barrier()
    ctx.timer_start(0)
    for _ in range(args.steps):
        step()
        phase.append(ctx.timings())          # waits for this step's last event (the step already synchronises once)
    ms_total = ctx.timer_stop(0)
    barrier()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    info = ctx.step_info()
    t = torch.tensor([ms_total, float(info["n_interactions"]), float(n), float(np.mean([p["walk_p2p"] for p in phase]))],
                     dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total, nint_total, n_total, walk_ms = float(tmax[0]), float(tsum[1]), float(tsum[2]), float(tmax[3])
    else:
        ms_total, nint_total, n_total, walk_ms = [float(x) for x in t]
    ms_step = ms_total / args.steps
    # per rank: step time and the rank's own phases (the job's step is the slowest rank's)
    per_rank = None
    if world > 1:
        mine = torch.tensor([float(t[0]) / args.steps, float(info["n_interactions"]), float(n)] +
                            [float(np.mean([p[k] for p in phase])) for k in ("tree", "frontier", "walk_p2p", "let")], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [float(x[0]) for x in allr], "interactions": [float(x[1]) for x in allr], "particles": [float(x[2]) for x in allr],
                    "tree_ms": [float(x[3]) for x in allr], "frontier_ms": [float(x[4]) for x in allr], "walk_p2p_ms": [float(x[5]) for x in allr],
                    "let_exposed_ms": [float(x[6]) for x in allr]}
    # Newton's third law over the whole periodic set: |sum a| / sum |a| (all ranks), a size-independent property check
    mom = torch.cat([acc.sum(0), torch.linalg.vector_norm(acc, dim=1).sum().reshape(1)])
    if world > 1:
        dist.all_reduce(mom, op=dist.ReduceOp.SUM)
    momentum_residual = float(torch.linalg.vector_norm(mom[:3]) / mom[3])

    # ---- e2e: the same step through the host-pointer C-ABI call: pinned host positions in, accelerations out ----
    e2e = None
    if not args.no_e2e:
        hpos = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
        hpos.copy_(pos)
        hacc = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
        lib = pn2gpu.lib()
        import ctypes as C

        def estep():
            pn2gpu._ck(lib.pn2_force_step(ctx.h, hpos.data_ptr(), 24, n, C.byref(dom), hacc.data_ptr(), 24))
        estep()
        barrier()
        ctx.timer_start(1)
        for _ in range(args.steps):
            estep()
        e_ms = ctx.timer_stop(1)
        barrier()
        te = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te[0]) / args.steps
        m_chk = min(n, 1 << 22)
        e2e = {"value": n_total / (e_ms * 1e-3), "unit": "particles/s", "h2d_bytes_per_step": int(n_total) * 24,
               "d2h_bytes_per_step": int(n_total) * 24, "ms_per_step": e_ms, "host_numa_binding_rank0": numa,
               # the host-pointer call must return what the device-resident step computed (same deterministic kernels)
               "max_abs_diff_vs_device_step_rank0": float((hacc[:m_chk].to(dev) - acc[:m_chk]).abs().max()),
               "rms_acc_rank0": float(torch.sqrt((acc[:m_chk] ** 2).sum(1).mean()))}

    # ---- PM long-range force on the same particles (not part of the metric) ----
    pm = None
    if args.pm and args.ic != "merger":
        apm = torch.empty_like(pos)
        for _ in range(2):
            ctx.pm_force_device(pos.data_ptr(), n, nside, apm.data_ptr())
        barrier()
        ctx.timer_start(2)
        for _ in range(3):
            ctx.pm_force_device(pos.data_ptr(), n, nside, apm.data_ptr())
        pm_ms = ctx.timer_stop(2) / 3
        ph = ctx.pm_timings()
        tp = torch.tensor([pm_ms], dtype=torch.float64, device=dev)
        mp = torch.cat([apm.sum(0), torch.linalg.vector_norm(apm, dim=1).sum().reshape(1)])
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            dist.all_reduce(mp, op=dist.ReduceOp.SUM)
        pm = {"ms": float(tp[0]), "nside": nside, "mesh_bytes": 8 * nside ** 3, "phases_ms_rank0": {k: float(v) for k, v in ph.items()},
              "momentum_residual": float(torch.linalg.vector_norm(mp[:3]) / mp[3]),
              "what": "CIC deposit + ncclAllReduce of the mesh + cuFFT D2Z / Green function / Z2D + 4-point gradient and CIC gather, particles in caller order"}
        del apm

    # ---- the reference's load-balance loop (src/photoNs.c:270-283, src/domains.c:21-160, 268-375) on the measured load ----
    rebalance = None
    if args.rebalance > 0 and world > 1:
        rebalance = []
        cur_pos, cur_n, cur_dom = pos, n, dom
        for it in range(args.rebalance + 1):
            if it > 0:
                # DTIME_THIS_DOMAIN = idxP2P + idxM2L (src/fmm.c:1069); DTIME_FRACTION = this * P / total (src/photoNs.c:281)
                inf = ctx.step_info()
                ld = torch.zeros(world, dtype=torch.float64, device=dev)
                ld[rank] = float(inf["n_p2p_pairs"] + inf["n_m2l_pairs"])
                dist.all_reduce(ld)
                load = (ld * world / (ld.sum() + 0.0001)).cpu().numpy()
                dtree.update(load)                                       # determine_split_domtree
                doms = dtree.boxes()
                ctx.set_comm(rank, world, doms, None)                    # same communicator, new domain table
                ptr, n_new = ctx.migrate_device(cur_pos.data_ptr(), 3, cur_n, dtree.splits)
                new_pos = torch.empty((n_new, 3), dtype=torch.float64, device=dev)
                pn2gpu._ck(pn2gpu.lib().pn2_migrate_fetch(ctx.h, new_pos.data_ptr()))
                cur_pos, cur_n, cur_dom = new_pos, n_new, doms[rank]
            cur_acc = torch.empty_like(cur_pos)
            barrier()
            ctx.timer_start(3)
            ctx.force_step_device(cur_pos.data_ptr(), cur_n, cur_acc.data_ptr(), cur_dom)
            ms_it = ctx.timer_stop(3)
            inf = ctx.step_info()
            v = torch.tensor([ms_it, float(inf["n_p2p_pairs"] + inf["n_m2l_pairs"]), float(cur_n), float(inf["n_interactions"])], dtype=torch.float64, device=dev)
            allv = [torch.zeros_like(v) for _ in range(world)]
            dist.all_gather(allv, v)
            A = torch.stack(allv).cpu().numpy()
            rebalance.append({"iter": it, "ms_max": float(A[:, 0].max()), "ms_mean": float(A[:, 0].mean()),
                              "imbalance_pct": float(100.0 * (1.0 - A[:, 1].sum() / (world * A[:, 1].max()))),       # src/photoNs.c:284
                              "load_max_over_mean": float(A[:, 1].max() / A[:, 1].mean()), "n_min": int(A[:, 2].min()), "n_max": int(A[:, 2].max()),
                              "interactions_total": float(A[:, 3].sum())})
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pps = n_total / (ms_step * 1e-3)
    ach = OPS_PER_INTERACTION * (nint_total / world) / (walk_ms * 1e-3)     # per GPU, dominant kernel
    fp32 = precision == pn2gpu.FP32
    nominal = 148 * (128 if fp32 else 64) * (clocks["sm_mhz"] or 1965.0) * 1e6 if clocks else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            # ncu capture of a smaller launch: scale the measured DRAM bytes by the launch's interaction count
            key = "walk_fused" if fp32 else "walk_fused_f64"
            traffic = tj[key + "_dram_bytes_per_launch"] / tj.get(key + "_interactions_per_launch", tj.get("interactions_per_launch")) * (nint_total / world)
        except Exception:
            traffic = None
    out = {"metric": METRIC, "value": pps, "unit": "particles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32" if precision == pn2gpu.FP32 else "f64", "data": "synthetic",
           "config": workload_config(args, world),
           "p2p_ginteractions_per_s": nint_total / (ms_step * 1e-3) / 1e9,
           "interactions_per_particle": nint_total / n_total,
           "migrate": migrate, "momentum_residual": momentum_residual, "rebalance": rebalance, "pm_long_range": pm, "per_rank": per_rank,
           "phases_ms": {k: float(np.mean([p[k] for p in phase])) for k in ("tree", "upward", "frontier", "walk_p2p", "m2l", "m2l_kernel", "downward", "let", "total")},
           "tree": {"nleaf": info["nleaf"], "nnode": info["nnode"], "levels": info["nlevel"], "m2l_pairs": info["n_m2l_pairs"],
                    "p2p_leaf_pairs": info["n_p2p_pairs"], "walk_visits": info["n_walk_visits"],
                    "frontier_bytes": info["frontier_bytes"]},
           "roofline": {"bound": "fma_pipe", "kernel": ("walk_fused_kernel" if fp32 else "walk_fused_f64_kernel") + " (list walk + P2P)",
                        "achieved": ach / 1e12, "peak": fma_peak / 1e12,
                        "unit": "Tops/s (FFMA/FMUL/FADD issue slots)" if fp32 else "Tops/s (DFMA/DMUL/DADD issue slots)", "frac": ach / fma_peak,
                        "peak_source": "measured in this run: pn2_fma_peak (independent %s chains, CUDA events)" % ("FFMA" if fp32 else "DFMA"),
                        "nominal_peak_at_sampled_clock": (nominal / 1e12) if nominal else None,
                        "ops_per_interaction": OPS_PER_INTERACTION, "ops_executed_per_interaction": 22 if fp32 else 28,
                        "frac_executed": (22 if fp32 else 28) / OPS_PER_INTERACTION * ach / fma_peak, "traffic": traffic,
                        "hbm_peak_gbs_measured": peaks.get("hbm_gbs")},
           "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e}
    # padded tiles: the kernel evaluates SW x SW lane products per leaf pair whatever the leaves hold
    sw = 8 if args.maxleaf <= 8 else (16 if args.maxleaf <= 16 else 32)
    lane_eff = info["n_interactions"] / max(1, info["n_p2p_pairs"] * sw * sw)
    out["tiles"] = {"slots": sw, "particles_per_leaf": info["n"] / max(1, info["nleaf"]), "lane_efficiency_rank0": lane_eff,
                    "note": "roofline.frac counts listed interactions; frac / lane_efficiency = what the lanes executed incl. padding slots"}
    out["roofline"]["frac_incl_padding_lanes"] = out["roofline"]["frac"] / lane_eff if lane_eff > 0 else None
    # the whole force step (tree, lists, LET, operators included) against the same roof: north_star's ">= 50 % at 512^3 on 8 GPUs"
    step_ops = OPS_PER_INTERACTION * nint_total / (ms_step * 1e-3) / world
    out["roofline"]["whole_step"] = {"achieved": step_ops / 1e12, "frac": step_ops / fma_peak,
                                     "frac_of_nominal_peak": (step_ops / nominal) if nominal else None,
                                     "note": "24 x all interactions / (GPUs x step time), max over ranks; peak as above"}
    m2l_ms = float(np.mean([p["m2l_kernel"] for p in phase]))
    if info["n_m2l_pairs"] > 0 and m2l_ms > 0:
        m2l_ops = 160.0 * info["n_m2l_pairs"] / (m2l_ms * 1e-3)
        out["m2l"] = {"pairs_rank0": info["n_m2l_pairs"], "kernel_ms": m2l_ms, "ops_per_pair": 160, "achieved_tops": m2l_ops / 1e12,
                      "dfma_peak_tops": dfma_peak / 1e12, "frac_of_dfma_peak": m2l_ops / dfma_peak,
                      "kernel": "m2l_warp_kernel (FP64; erfc / exp from degree-10 tables, 1/r from MUFU.RSQ64H + Newton: no libm in the long/short build)"}
    if not args.no_cpu_baseline and world == 1 and args.ic == "lcdm":
        cores = host_cores()
        nr = 1
        while nr * 2 <= cores:
            nr *= 2
        try:
            r = reference_cpu_run(args, args.cpu_sample_side, nr, keep=True)
            out["cpu_baseline"] = {"value": r["pps"], "unit": "particles/s", "cores": r["cores"], "kind": r["kind"],
                                   "sample": f"{r['side']}^3 particles of the same generator, NSIDE {r['nside']}, one force evaluation, "
                                             f"NP={r['cores']} ranks of the unmodified reference (PM excluded), "
                                             f"{r['interactions_per_particle']:.0f} interactions/particle, {r['sec']:.1f} s",
                                   "p2p_ginteractions_per_s": r["ips"] / 1e9}
            if "kept" in r:
                try:
                    out["parity"] = parity_vs_reference(args, r)
                except Exception as ex:
                    out["parity"] = {"pass": False, "error": repr(ex)[:300]}
        except Exception as ex:  # the bench line must still be printed
            out["cpu_baseline"] = {"value": None, "unit": "particles/s", "cores": cores, "kind": "unavailable", "sample": repr(ex)[:300]}
    print(json.dumps(out), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
