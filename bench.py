#!/usr/bin/env python
"""bench.py -- the hot path of efliks/pcetk on B200, measured on the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lysozyme|ifp20|defensin|synthM|synthA|synthA36] [--impl b200|reference]

Headline (default) workload = BASELINE.json configs[1]: the lysozyme fixture (21 sites / 44 instances / 3 double-move pairs),
Metropolis MC titration curve pH 0..20 step 0.5 (41 points), 500 equilibration + 20000 production scans per chain
(the reference's defaults), `--chains` chains per pH value per GPU.  One "step" = one whole titration curve.
metric = MC trial moves per second (whole job).  Prints ONE JSON line (rank 0).

  value : device-resident model, kernel + (N>1) the one NCCL all-reduce of the occupancy counts; CUDA events.
  e2e   : the same curve through the public API from HOST arrays: the model is re-loaded, symmetrised and uploaded, the
          kernel runs, and counts come back to the host, every step; wall clock around a device synchronise.  One GPU:
          CEModel.CalculateProbabilitiesBatch; several: pcetk_b200.distributed (chains / chunk range sharded, ONE all-reduce).
  roofline : the resource that bounds the dominant kernel, measured in the same run: issue slots (chain-per-thread and small
          chain-per-warp Monte Carlo: warp-instructions per trial from the committed ncu digest x trials/s, against an
          issue-rate probe), L2 read bandwidth (1000-site Monte Carlo: bytes from the kernel's own move counters, against a
          streaming-read probe), FP64 (enumeration: one DFMA per state and role rotation, against a DFMA probe).
  workloads : the other BASELINE configs in the same invocation (short runs): synthA = 2^32-state analytic enumeration
          (states/s; strong scaling over GPUs), synthM = 1000 sites / 3000 instances Monte Carlo (a slice of the
          65536 chains x 41 pH shape: `chains` per pH value per GPU, nequil 500), ifp20, defensin.  Each carries its own
          ms_per_step, e2e, roofline and (one GPU) cpu_baseline.  `--workload X` makes X the headline instead and skips the rest.
  --impl reference : the reference's own C (oracle/_ref, compiled unmodified from /root/reference in the build
          container) on all host threads, one independent single-chain curve per thread per step.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def ph_grid():
    return np.arange(0.0, 20.0 + 1e-9, 0.5)          # 41 points, TitrationCurves.py:62 with curveStop=20


def load_record(workload):
    from pcetk_b200.formats import ModelRecord
    from pcetk_b200 import synthetic

    if workload in ("lysozyme", "ifp20", "defensin"):
        return ModelRecord.from_npz_dict(np.load(os.path.join(ROOT, "tests", "golden", workload + ".npz")))
    if workload == "synthM":
        return synthetic.synth_m(1000)
    if workload == "synthA":
        return synthetic.synth_a(32)
    if workload == "synthA36":
        return synthetic.synth_a(36)
    raise SystemExit("unknown workload " + workload)


class ClockSampler(object):
    """nvidia-smi clocks during the timed region (B200_PROFILING.md, the clocks line)."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])); smax.append(float(t[2]))
            except ValueError:
                continue
            for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arms (the only place outside tests/ and smoke() that executes oracle/)
# ---------------------------------------------------------------------------------------------------
def cpu_curve_worker(record, nequil, nprod, seed, grid, out, index, use_ref):
    """One single-chain titration curve on one host thread; returns trial moves done."""
    import oracle

    if use_ref:
        ref = oracle.Ref(record)
        mc = ref.mc_create(2.0, nequil, nprod, seed)
        npairs = len(ref.mc_pairs(mc)[0])
        for ph in grid:
            ref.mc_run(mc, float(ph))
        ref.close()
    else:
        port = oracle.Port()
        wsym = port.symmetrize(record.interactions)
        pairs = port.find_pairs(record, wsym, 2.0)[:2]
        npairs = len(pairs[0])
        port.mc_mt(record, wsym, pairs, nequil, nprod, seed, grid)
    out[index] = len(grid) * (nequil + nprod) * (record.nsites + npairs)


def cpu_baseline(record, nequil, nprod, grid, threads, repeats=1):
    import oracle

    use_ref = oracle.have_ref()
    if not use_ref:
        oracle.build()
    moves = [0] * threads
    start = time.perf_counter()
    for rep in range(repeats):
        workers = [threading.Thread(target=cpu_curve_worker, args=(record, nequil, nprod, 1000 + 17 * k + rep, grid, moves, k, use_ref)) for k in range(threads)]
        for w in workers:
            w.start()
        for w in workers:
            w.join()
    elapsed = time.perf_counter() - start
    total = sum(moves) * repeats
    return total / elapsed, elapsed, ("reference" if use_ref else "port")


def cpu_baseline_analytic(nsites=24, ph=7.0):
    """Exhaustive enumeration of the reference (EnergyModel_CalculateProbabilitiesAnalytically) on one host core for the
    `nsites`-site variant of the synth-A generator: the reference cannot run 2^32 states (2^26 cap, a 32 GiB array), so
    the largest case that stays within ~10 s is timed.  Returns (states/s, seconds, kind, nstates)."""
    import oracle
    from pcetk_b200 import synthetic

    record = synthetic.synth_a(nsites)
    use_ref = oracle.have_ref()
    start = time.perf_counter()
    if use_ref:
        ref = oracle.Ref(record)
        ref.analytic(float(ph))
        ref.close()
    else:
        oracle.build()
        port = oracle.Port()
        port.analytic(record, port.symmetrize(record.interactions), float(ph))
    elapsed = time.perf_counter() - start
    return record.nstates / elapsed, elapsed, ("reference" if use_ref else "port"), record.nstates


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    record = load_record(args.workload if args.workload in ("lysozyme", "ifp20", "defensin") else "lysozyme")
    grid = ph_grid()
    threads = args.cpu_threads or (os.cpu_count() or 1)
    nequil, nprod = 500, args.cpu_nprod
    for _ in range(max(args.warmup, 0) and 1):
        cpu_baseline(record, 50, 500, grid[:4], threads)
    t0 = time.perf_counter()
    rate, elapsed, kind = cpu_baseline(record, nequil, nprod, grid, threads, repeats=args.steps)
    value = rate
    sample = "%d host threads x %d step(s) x one single-chain %s curve (41 pH, %d+%d scans, MT19937)" % (threads, args.steps, record.name, nequil, nprod)
    line = {
        "impl": "reference", "metric": "mc_trial_moves_per_s", "value": value, "unit": "moves/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * elapsed / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "reference fixture (tests/lysozyme results.tgz), synthetic chains",
        "config": {"workload": "%s MC titration curve pH 0..20 step 0.5 (41 pts), nequil %d, nprod %d, 1 chain per thread" % (record.name, nequil, nprod)},
        "cpu_baseline": {"value": value, "unit": "moves/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
MC_DEFAULTS = {   # chains per pH per GPU (0 = whole waves of the saturating launch geometry), nprod, nequil
    "lysozyme": (0, 20000, 500), "defensin": (0, 20000, 500), "ifp20": (256, 2000, 500), "synthM": (0, 100, 500)}


class Context(object):
    """Device, process group, launch stream and the measured peaks of this run."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import pcetk_b200 as P
        from pcetk_b200 import _native

        self.torch, self.dist, self.P, self.native = torch, dist, P, _native
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.device)
        self.stream = torch.cuda.current_stream()
        self.lib = _native.library()
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.device)      # > 126 MB L2
        self.peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as handle:
                self.peaks = json.load(handle)
        except (OSError, ValueError):
            pass
        self.digests = {}
        try:
            with open(os.path.join(ROOT, "profiles", "r02_summary.json")) as handle:
                self.digests = json.load(handle)
        except (OSError, ValueError):
            pass
        self._probes = {}

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, value):
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def probe(self, kind):
        """Measured on-chip peaks of this device: 'dfma' (1e9 DFMA/s), 'shared' (GB/s), 'issue' (1e9 warp-instr/s), 'l2' (GB/s)."""
        import ctypes as C
        if kind not in self._probes:
            value = C.c_double(0.0)
            if kind == "l2":
                self.native.check(self.lib.pce_probe_read_bandwidth(self.local_rank, 48 * 1024 * 1024, 40, C.byref(value)))
            else:
                self.native.check(self.lib.pce_probe_on_chip(self.local_rank, {"dfma": 0, "shared": 1, "issue": 2}[kind], C.byref(value)))
            self._probes[kind] = value.value
        return self._probes[kind]

    def timed(self, step, steps, warmup, clocks=False):
        """W untimed steps, then K steps timed with CUDA events on the launch stream (L2 flushed between steps, outside the
        events), barrier + synchronise on both sides, max over ranks.  -> (ms of all K steps, clock record, kernel launches)"""
        torch = self.torch
        for _ in range(warmup):
            step()
        self.barrier()
        self.lib.pce_launch_count(1)
        sampler = ClockSampler(self.local_rank) if clocks and self.rank == 0 else None
        if sampler:
            sampler.start()
        events = []
        for _ in range(steps):
            self.flush.fill_(1)
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start.record(self.stream)
            step()
            stop.record(self.stream)
            events.append((start, stop))
        self.barrier()
        clock_info = sampler.stop() if sampler else None
        launches = int(self.lib.pce_launch_count(0))
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in events)), clock_info, launches

    def timed_wall(self, step, steps, warmup):
        for _ in range(warmup):
            step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.torch.cuda.synchronize()
        return self.max_over_ranks(time.perf_counter() - t0)


def model_h2d_bytes(record, nph):
    return record.ninstances * ((record.ninstances + 3) // 4 * 4) * 8 + record.ninstances * 20 + record.nsites * 8 + nph * 8


def mc_workload(ctx, name, steps, warmup, chains=0, nprod=0, nequil=-1, waves=2, kernel=0, e2e=True, cpu=True, headline=False):
    """Monte Carlo titration curve (41 pH values) of one model -> dict of the bench-line keys of this workload."""
    import ctypes as C
    P, lib, native, torch = ctx.P, ctx.lib, ctx.native, ctx.torch
    record = load_record(name)
    model = P.MEADModel(record, log=None, device=ctx.local_rank)
    em = model.energyModel
    em.SetStream(ctx.stream.cuda_stream)
    default_chains, default_nprod, default_nequil = MC_DEFAULTS[name]
    chains = chains or default_chains
    nprod = nprod or default_nprod
    nequil = nequil if nequil >= 0 else default_nequil
    grid = ph_grid()
    nph = len(grid)
    sampler = P.MCModelDefault(nequil=nequil, nprod=nprod, randomSeed=20241019, nchains=max(chains, 1), kernel=kernel)
    model.DefineMCModel(sampler, log=None)
    if not chains:
        sampler.nchains = 1 << 20          # the launch geometry of a saturating ensemble (the planner sizes blocks by the ensemble)
        plan = sampler.LaunchPlan(nph)
        chains = max(1, (waves * plan["chains_per_wave"]) // nph)
        sampler.nchains = chains
    plan = sampler.LaunchPlan(nph)
    sampler.chainOffset = ctx.rank * chains
    sampler._configure()
    npairs = sampler.npairs
    nmoves = record.nsites + npairs
    d_counts = torch.zeros((nph, record.ninstances), dtype=torch.int64, device=ctx.device)
    d_counters = torch.zeros((nph, 4), dtype=torch.int64, device=ctx.device)
    em.Upload()

    def step_device():
        # occupancy counts and move counters; no energy sums, like the reference's quiet path (MCModelDefault.pyx:60-78)
        native.check(lib.pce_mc_run_device(sampler._handle, grid, nph, 0, ctx.rank * chains, C.c_void_p(d_counts.data_ptr()),
                                           C.c_void_p(d_counters.data_ptr()), None))
        if ctx.world > 1:
            ctx.dist.all_reduce(d_counts)            # the path's only exchange: per-(pH, instance) occupancy counts

    def step_e2e():
        em.Load(gintr=record.gintr, gmodel=record.gmodel, protons=record.protons, interactions=record.interactions)
        em.SymmetrizeInteractions(log=None)
        if ctx.world > 1:
            from pcetk_b200.distributed import mc_probabilities
            return mc_probabilities(sampler, grid, total_chains=chains * ctx.world, rank=ctx.rank, world=ctx.world, device=ctx.device, phIndexOffset=0)[0]
        return model.CalculateProbabilitiesBatch(grid)

    units_per_step = float(nph) * chains * (nequil + nprod) * nmoves * ctx.world
    elapsed_ms, clock_info, launches = ctx.timed(step_device, steps, warmup, clocks=headline)
    value = units_per_step * steps / (elapsed_ms * 1e-3)
    out = {
        "metric": "mc_trial_moves_per_s", "value": value, "unit": "moves/s", "steps": steps, "warmup": warmup, "ms_per_step": elapsed_ms / max(steps, 1),
        "scaling": "weak", "dtype": "f64",
        "data": "reference fixture model, synthetic chains" if name in ("lysozyme", "ifp20", "defensin") else "synthetic",
        "config": {"workload": "%s (%d sites / %d instances / %d pairs) MC titration curve pH 0..20 step 0.5 (41 pts), nequil %d, nprod %d, %d chains per pH per GPU" % (
                       record.name, record.nsites, record.ninstances, npairs, nequil, nprod, chains),
                   "l2": "flushed between timed steps (256 MiB write)",
                   "parallelism": "chains sharded over %d GPU(s), one all-reduce of counts" % ctx.world, "launch": plan},
        "gpu_launches": launches,
    }
    if name == "synthM":
        out["config"]["shape"] = ("slice of BASELINE configs[4] (65536 chains x 41 pH, nequil 500): %d chains per pH value per GPU, the same per-chain work "
                                  "(500 + %d scans); the full shape on 8 GPUs is recorded in profiles/r02_synthM_full_shape.json" % (chains, nprod))
    if clock_info is not None:
        out["clocks"] = clock_info
    c = d_counters.sum(dim=0).cpu().numpy().astype(float)       # production trials of this rank: moves, accepted, double moves, accepted
    out["config"]["acceptance"] = {"single": c[1] / max(c[0], 1.0), "double": c[3] / max(c[2], 1.0), "overall": (c[1] + c[3]) / max(c[0] + c[2], 1.0)}
    if e2e:
        e2e_s = ctx.timed_wall(step_e2e, steps, min(warmup, 2))
        out["e2e"] = {"value": units_per_step * steps / e2e_s, "unit": "moves/s", "h2d_bytes_per_step": int(model_h2d_bytes(record, nph)),
                      "d2h_bytes_per_step": int(nph * (record.ninstances * 8 + 4 * 8 + 8))}
    if ctx.rank == 0:
        out["roofline"] = mc_roofline(ctx, record, plan, name, value / ctx.world, out["ms_per_step"], c, chains, nequil, nprod, npairs)
    if ctx.rank == 0 and ctx.world == 1 and e2e and headline and name in ("lysozyme", "ifp20", "defensin"):
        # BASELINE's third metric: wall time of ONE titration curve through the public API (TitrationCurves), host
        # arrays in, curves out.  (a) the reference's statistical effort per pH value (20000 production scans) split
        # over 64 chains; (b) the package defaults (64 chains x 20000 scans = 64x the reference's samples).
        curve = {}
        for label, nch, npr in (("reference_effort", 64, 313), ("defaults", 64, 20000 if name != "ifp20" else 2000)):
            cm = P.MEADModel(record, log=None, device=ctx.local_rank)
            cm.DefineMCModel(P.MCModelDefault(nequil=500, nprod=npr, randomSeed=7, nchains=nch), log=None)
            P.TitrationCurves(cm, log=None, curveSampling=0.5, curveStart=0.0, curveStop=20.0).CalculateCurves(log=None)     # warm-up
            torch.cuda.synchronize()
            walls = []
            for _ in range(3):        # median of three curves (each from host arrays to half-pK values)
                t0 = time.perf_counter()
                tc = P.TitrationCurves(cm, log=None, curveSampling=0.5, curveStart=0.0, curveStop=20.0)
                tc.CalculateCurves(log=None)
                tc.CalculateHalfpKs()
                walls.append(1000.0 * (time.perf_counter() - t0))
            curve[label] = {"wall_ms": sorted(walls)[1], "wall_ms_runs": walls, "chains": nch, "nequil": 500, "nprod_per_chain": npr, "ph_points": tc.nsteps,
                            "launch": cm.sampler.LaunchPlan(tc.nsteps) if hasattr(cm, "sampler") and cm.sampler is not None else None}
        out["titration_curve"] = curve
    if ctx.rank == 0 and ctx.world == 1 and cpu:
        if name == "synthM":
            # 20 + 100 scans of one chain at pH 7: ~8e5 trial moves, 5-10 s on one host core
            rate, seconds, kind = cpu_baseline(record, 20, 100, grid[14:15], 1)
            sample = "one single-chain run at pH 7 (20 + 100 scans of %d trial moves, MT19937), %.1f s on one host core" % (nmoves, seconds)
        elif name == "ifp20":
            rate, seconds, kind = cpu_baseline(record, 500, 20000, grid[10:20], 1)
            sample = "one single-chain run over 10 pH values (pH 5..9.5, 500+20000 scans, MT19937), %.1f s on one host core" % seconds
        else:
            ncurves = 6 if headline else 2
            rate, seconds, kind = cpu_baseline(record, 500, 20000, grid, 1, repeats=ncurves)
            sample = "%d single-chain %s curve(s) (41 pH, 500+20000 scans, MT19937), %.1f s on one host core" % (ncurves, record.name, seconds)
            if "titration_curve" in out:
                out["titration_curve"]["reference_one_core_wall_ms"] = 1000.0 * seconds / ncurves
        out["cpu_baseline"] = {"value": rate, "unit": "moves/s", "cores": 1, "kind": kind, "sample": sample}
    return out


def mc_roofline(ctx, record, plan, name, value_per_gpu, ms_per_step, counters, chains, nequil, nprod, npairs):
    """The binding on-chip resource of the Monte Carlo kernel that ran, with a denominator measured in this run."""
    nmoves = record.nsites + npairs
    reference_bytes = (record.nsites * 20.0 * record.nsites + npairs * 36.0 * record.nsites) / nmoves
    common = {"kernel": "k_mc_thread" if plan["kernel"] == 1 else "k_mc_warp", "kernel_ms_per_launch": ms_per_step,
              "hbm": "not the limiter: the working set is on chip (shared memory / L2); DRAM traffic per launch in profiles/",
              "reference_algorithm_bytes_per_unit": reference_bytes, "reference_algorithm_effective_gbs": reference_bytes * value_per_gpu / 1e9}
    if name == "synthM":
        # chain-per-warp kernel on the 1000-site model: difference rows stream from L2.  Algorithmic L2 bytes of one launch = one
        # difference row per changed site of every accepted move + one cross-term double per double-move proposal + the rows read
        # by the field initialisation; accepted moves are counted in production scans and scaled to all scans.
        nph = len(ph_grid())
        scale = float(nequil + nprod) / float(nprod)
        row_bytes = (record.ninstances - record.nsites) * 8.0
        l2_bytes = scale * ((counters[1] + 2.0 * counters[3]) * row_bytes + counters[2] * 8.0) + float(nph) * chains * record.nsites * row_bytes
        peak = ctx.probe("l2")
        achieved = l2_bytes / (ms_per_step * 1e-3) / 1e9
        return dict(common, bound="l2", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / max(peak, 1e-9), traffic=l2_bytes,
                    peak_source="streaming 128-bit reads of an L2-resident 48 MiB buffer, measured in this run (pce_probe_read_bandwidth)",
                    note="achieved = algorithmic L2 bytes from the kernel's own move counters / launch time; the kernel's limiter next to L2 is "
                         "the L1TEX wavefront rate (profiles/README.md)")
    digest = ctx.digests.get("%s_%s" % ("k_mc_thread" if plan["kernel"] == 1 else "k_mc_warp", name), None)
    if digest is None or not digest.get("warp_instructions_per_trial"):
        return dict(common, bound="issue", achieved=None, peak=ctx.probe("issue"), unit="1e9 warp-instructions/s", frac=None, traffic=None,
                    note="no ncu digest of this kernel build under profiles/ (r02_summary.json): instructions per trial unknown")
    per_trial = float(digest["warp_instructions_per_trial"])
    peak = ctx.probe("issue")
    achieved = value_per_gpu * per_trial / 1e9
    return dict(common, bound="issue", achieved=achieved, peak=peak, unit="1e9 warp-instructions/s", frac=achieved / max(peak, 1e-9),
                traffic=digest.get("dram_bytes_per_launch"), warp_instructions_per_unit=per_trial, digest=digest.get("source"),
                peak_source="independent FP32 FMAs on every SM (128 lanes: one warp-instruction per scheduler and clock), measured in this run (pce_probe_on_chip kind 2); nominal 148 SMs x 4 per clock",
                note="achieved = trial moves/s of this run x warp-instructions per trial of the committed ncu digest of the same kernel")


def analytic_workload(ctx, name, steps, warmup, e2e=True, cpu=True):
    """Exhaustive enumeration over the 41-point pH grid: states/s, chunk range sharded over ranks (strong scaling)."""
    import ctypes as C
    P, lib, native, torch = ctx.P, ctx.lib, ctx.native, ctx.torch
    record = load_record(name)
    model = P.MEADModel(record, log=None, device=ctx.local_rank)
    em = model.energyModel
    em.SetStream(ctx.stream.cuda_stream)
    em.analyticStatesLimit = 2 ** 44
    grid = ph_grid()
    nbins = int(lib.pce_analytic_bins_size(em._handle))
    max_protons = nbins // (1 + record.ninstances) - 1
    groups = em._phGroups(grid, max_protons)
    bins = [torch.zeros(nbins, dtype=torch.float64, device=ctx.device) for _ in groups]
    em.Upload()

    def step_device():
        for group, buf in zip(groups, bins):
            sub = np.ascontiguousarray(grid[group])
            native.check(lib.pce_analytic_bins(em._handle, sub, len(sub), 0, ctx.rank, ctx.world, float(em.analyticStatesLimit), C.c_void_p(buf.data_ptr())))
            if ctx.world > 1:
                ctx.dist.all_reduce(buf)

    def step_e2e():
        em.Load(gintr=record.gintr, gmodel=record.gmodel, protons=record.protons, interactions=record.interactions)
        em.SymmetrizeInteractions(log=None)
        if ctx.world > 1:
            from pcetk_b200.distributed import analytic_probabilities
            return analytic_probabilities(model, grid, rank=ctx.rank, world=ctx.world, device=ctx.device)
        return model.CalculateProbabilitiesBatch(grid)

    nstates = float(record.nstates)
    elapsed_ms, _clock, launches = ctx.timed(step_device, steps, warmup)
    value = nstates * steps / (elapsed_ms * 1e-3)
    out = {
        "metric": "analytic_microstates_per_s", "value": value, "unit": "states/s", "steps": steps, "warmup": warmup, "ms_per_step": elapsed_ms / max(steps, 1),
        "scaling": "strong", "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic %d sites x 2 instances, 2^%d microstates, analytic probabilities of all instances on pH 0..20 step 0.5 (41 pts) per step" % (
                       record.nsites, record.nsites),
                   "l2": "flushed between timed steps (256 MiB write)", "enumeration_passes_per_curve": len(groups),
                   "parallelism": "chunk range sharded over %d GPU(s), one all-reduce of the proton-count bins" % ctx.world},
        "gpu_launches": launches,
    }
    if e2e:
        e2e_s = ctx.timed_wall(step_e2e, steps, min(warmup, 2))
        out["e2e"] = {"value": nstates * steps / e2e_s, "unit": "states/s", "h2d_bytes_per_step": int(model_h2d_bytes(record, len(grid))),
                      "d2h_bytes_per_step": int(nbins * 8 * len(groups))}
    if ctx.rank == 0:
        # one DFMA per state and role rotation (two rotations cover all sites): algorithmic flops = 4 per state
        peak = 2.0 * ctx.probe("dfma") / 1e3                     # TFLOP/s
        achieved = 4.0 * value / ctx.world / 1e12
        digest = ctx.digests.get("k_enum", {})
        out["roofline"] = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / max(peak, 1e-9),
                           "traffic": digest.get("dram_bytes_per_launch"), "kernel": "k_enum", "kernel_ms_per_launch": out["ms_per_step"] / (2.0 * len(groups)),
                           "peak_source": "8 independent DFMA chains per thread on every SM, measured in this run (pce_probe_on_chip kind 0)",
                           "shared_memory_peak_gbs": ctx.probe("shared"),
                           "note": "algorithmic flops = 2 role rotations x 1 DFMA per state; the kernel's own limiter is the shared-memory broadcast of "
                                   "the tile table (profiles/README.md); HBM traffic ~0 by construction",
                           "reference_algorithm_bytes_per_unit": record.nsites * (record.nsites - 1) / 2 * 8.0}
    if ctx.rank == 0 and ctx.world == 1 and cpu:
        # the reference cannot run this workload at all (2^26 cap, an nstates-sized array); its per-state cost is nsites (nsites - 1) / 2
        # gathers + one exp per pH value, so the 24-site rate is an UPPER bound of what it would do on 32 sites (276 against 496 gathers)
        rate, seconds, kind, _n = cpu_baseline_analytic(24, 7.0)
        out["cpu_baseline"] = {"value": rate, "unit": "states/s", "cores": 1, "kind": kind,
                               "sample": "24-site variant of the same generator, 2^24 states at ONE pH value, %.1f s on one host core; the "
                                         "reference caps exhaustive enumeration at 2^26 states and needs one pass per pH value "
                                         "(x 41 for this curve), so this curve extrapolates to >= %.0f core-hours" % (seconds, nstates * 41 / rate / 3600.0)}
    return out


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--workload", default="", choices=["", "lysozyme", "ifp20", "defensin", "synthM", "synthA", "synthA36"],
                        help="make this workload the headline and skip the side workloads (default: lysozyme + all side workloads)")
    parser.add_argument("--chains", type=int, default=0, help="chains per pH value per GPU (default depends on the workload)")
    parser.add_argument("--nprod", type=int, default=0, help="production scans per chain (default: reference's 20000; ifp20 2000; synthM 100)")
    parser.add_argument("--nequil", type=int, default=-1)
    parser.add_argument("--waves", type=int, default=2, help="when --chains is not given: chains = waves x (chains resident on the GPU) / 41")
    parser.add_argument("--kernel", type=int, default=0, help="MC kernel: 0 auto, 1 chain-per-thread, 2 chain-per-warp")
    parser.add_argument("--cpu-threads", type=int, default=0)
    parser.add_argument("--cpu-nprod", type=int, default=20000)
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-e2e", action="store_true")
    parser.add_argument("--no-side-workloads", action="store_true")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        if not args.workload:
            args.workload = "lysozyme"
        run_reference_arm(args)
        return

    ctx = Context()
    headline = args.workload or "lysozyme"
    e2e, cpu = not args.no_e2e, not args.no_cpu_baseline
    if headline.startswith("synthA"):
        result = analytic_workload(ctx, headline, args.steps, args.warmup, e2e=e2e, cpu=cpu)
    else:
        result = mc_workload(ctx, headline, args.steps, args.warmup, chains=args.chains, nprod=args.nprod, nequil=args.nequil, waves=args.waves,
                             kernel=args.kernel, e2e=e2e, cpu=cpu, headline=True)
    side = {}
    if not args.workload and not args.no_side_workloads:
        # the other BASELINE configs, short runs (W = 3 warm-up steps as the timing rules ask)
        side["synthA"] = analytic_workload(ctx, "synthA", 10, 3, e2e=e2e, cpu=cpu)
        # 16x the states (2^36): the strong-scaling curve whose fixed per-curve cost (reductions, the all-reduce) no longer shows
        side["synthA36"] = analytic_workload(ctx, "synthA36", 3, 3, e2e=False, cpu=False)
        side["synthM"] = mc_workload(ctx, "synthM", 2, 3, e2e=e2e, cpu=cpu)
        side["ifp20"] = mc_workload(ctx, "ifp20", 3, 3, e2e=e2e, cpu=cpu)
        side["defensin"] = mc_workload(ctx, "defensin", 2, 3, e2e=e2e, cpu=cpu)
    if ctx.rank == 0:
        line = {"metric": result["metric"], "value": result["value"], "unit": result["unit"], "n_gpus": ctx.world, "steps": result["steps"],
                "warmup": result["warmup"], "ms_per_step": result["ms_per_step"], "higher_is_better": True, "scaling": result["scaling"],
                "vs_baseline": None, "dtype": result["dtype"], "data": result["data"], "config": result["config"],
                "clocks": result.get("clocks"), "gpu_launches": result["gpu_launches"] + sum(w["gpu_launches"] for w in side.values()),
                "roofline": result.get("roofline")}
        for key in ("e2e", "titration_curve", "cpu_baseline"):
            if key in result:
                line[key] = result[key]
        if side:
            for w in side.values():
                w["n_gpus"] = ctx.world
                w["higher_is_better"] = True
            line["workloads"] = side
        line["measured_peaks"] = dict(ctx._probes, hbm_gbs=ctx.peaks.get("hbm_gbs"), source="probes of this run; hbm_gbs from MEASURED_PEAKS.json")
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
