#!/usr/bin/env python
"""bench.py -- the hot path of efliks/pcetk on B200, measured on the metric BASELINE.json names.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lysozyme|ifp20|synthM|synthA] [--impl b200|reference]

Default workload = BASELINE.json configs[1]: the lysozyme fixture (21 sites / 44 instances / 3 double-move pairs),
Metropolis MC titration curve pH 0..20 step 0.5 (41 points), 500 equilibration + 20000 production scans per chain
(the reference's defaults), `--chains` chains per pH value per GPU.  One "step" = one whole titration curve.
metric = MC trial moves per second (whole job).  Prints ONE JSON line (rank 0).

  value : device-resident model, kernel + (N>1) the one NCCL all-reduce of the occupancy counts; CUDA events.
  e2e   : the same curve through the public API (CEModel.CalculateProbabilitiesBatch) from HOST arrays: the model is
          re-loaded, symmetrised and uploaded, the kernel runs, and counts come back to the host, every step; wall clock
          around a device synchronise.
  --impl reference : the reference's own C (oracle/_ref, compiled unmodified from /root/reference in the build
          container) on all host threads, one independent single-chain curve per thread per step.

Other workloads (`ifp20`, `synthM` = 1000 sites MC, `synthA` = 2^32-state enumeration) report their own metric and are
extra evidence, not the headline line.  synthM defaults to two waves of resident chains (50 per pH value), 100 + 400 scans
(0.35 s per curve on one B200); BASELINE's full shape is `--chains 65536 --nequil 500 --nprod N` (2.7e6 chains: about
8 minutes per step on one GPU at 1.9e10 moves/s for N = 50).
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
def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--workload", default="lysozyme", choices=["lysozyme", "ifp20", "defensin", "synthM", "synthA"])
    parser.add_argument("--chains", type=int, default=0, help="chains per pH value per GPU (default depends on the workload)")
    parser.add_argument("--nprod", type=int, default=0, help="production scans per chain (default: reference's 20000; ifp20 2000; synthM 400)")
    parser.add_argument("--nequil", type=int, default=-1)
    parser.add_argument("--waves", type=int, default=2, help="when --chains is not given: chains = waves x (chains resident on the GPU) / 41")
    parser.add_argument("--kernel", type=int, default=0, help="MC kernel: 0 auto, 1 chain-per-thread, 2 chain-per-warp")
    parser.add_argument("--cpu-threads", type=int, default=0)
    parser.add_argument("--cpu-nprod", type=int, default=20000)
    parser.add_argument("--no-cpu-baseline", action="store_true")
    parser.add_argument("--no-e2e", action="store_true")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import pcetk_b200 as P
    from pcetk_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    record = load_record(args.workload)
    model = P.MEADModel(record, log=None, device=local_rank)
    em = model.energyModel
    stream = torch.cuda.current_stream()
    em.SetStream(stream.cuda_stream)
    lib = _native.library()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)      # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    result = {}
    if args.workload == "synthA":
        # analytic enumeration: states/s over the 41-point grid, chunk range sharded over ranks
        import ctypes as C
        em.analyticStatesLimit = 2 ** 40
        grid_all = ph_grid()
        max_protons = int(lib.pce_analytic_bins_size(em._handle) // (1 + record.ninstances)) - 1
        groups = em._phGroups(grid_all, max_protons)
        nbins = lib.pce_analytic_bins_size(em._handle)
        bins = [torch.zeros(nbins, dtype=torch.float64, device=device) for _ in groups]
        em.Upload()

        def step_device():
            for group, buf in zip(groups, bins):
                sub = np.ascontiguousarray(grid_all[group])
                _native.check(lib.pce_analytic_bins(em._handle, sub, len(sub), 0, rank, world, 0.0, C.c_void_p(buf.data_ptr())))
                if world > 1:
                    dist.all_reduce(buf)

        def step_e2e():
            em.Load(gintr=record.gintr, gmodel=record.gmodel, protons=record.protons, interactions=record.interactions)
            em.SymmetrizeInteractions(log=None)
            return model.CalculateProbabilitiesBatch(grid_all)

        units_per_step = float(record.nstates)          # microstates enumerated per curve (all 41 pH values served)
        metric, unit = "analytic_microstates_per_s", "states/s"
        launch_info = {"passes": len(groups)}
        workload = "synthetic 32 sites x 2 instances, 2^32 microstates, analytic probabilities on pH 0..20 step 0.5 (41 pts) per step"
        h2d = record.ninstances * ((record.ninstances + 3) // 4 * 4) * 8 + record.ninstances * 20 + record.nsites * 8
        d2h = int(nbins) * 8 * len(groups)
        algo_bytes_per_unit = 0.0
    else:
        defaults = {"lysozyme": (1024, 20000, 500), "defensin": (1024, 20000, 500), "ifp20": (256, 2000, 500), "synthM": (50, 400, 100)}
        chains, nprod, nequil = defaults[args.workload]
        chains = args.chains or chains
        nprod = args.nprod or nprod
        nequil = args.nequil if args.nequil >= 0 else nequil
        grid_all = ph_grid()
        sampler = P.MCModelDefault(nequil=nequil, nprod=nprod, randomSeed=20241019, nchains=chains, kernel=args.kernel)
        model.DefineMCModel(sampler, log=None)
        if not args.chains:
            # size the ensemble to whole waves of the launch geometry (chain-major linear grid)
            sampler.nchains = 1 << 20          # the launch geometry of a saturating ensemble (the planner sizes blocks by the ensemble)
            plan = sampler.LaunchPlan(len(grid_all))
            chains = max(1, (args.waves * plan["chains_per_wave"]) // len(grid_all))
            sampler.nchains = chains
        plan = sampler.LaunchPlan(len(grid_all))
        sampler.chainOffset = rank * chains
        npairs = sampler.npairs
        nmoves = record.nsites + npairs
        sampler._configure()
        d_counts = torch.zeros((len(grid_all), record.ninstances), dtype=torch.int64, device=device)
        d_counters = torch.zeros((len(grid_all), 4), dtype=torch.int64, device=device)
        d_esum = torch.zeros(len(grid_all), dtype=torch.float64, device=device)
        em.Upload()
        import ctypes as C

        def step_device():
            _native.check(lib.pce_mc_run_device(sampler._handle, grid_all, len(grid_all), 0, rank * chains, C.c_void_p(d_counts.data_ptr()),
                                                C.c_void_p(d_counters.data_ptr()), C.c_void_p(d_esum.data_ptr())))
            if world > 1:
                dist.all_reduce(d_counts)            # the path's only exchange: per-(pH, instance) occupancy counts

        def step_e2e():
            em.Load(gintr=record.gintr, gmodel=record.gmodel, protons=record.protons, interactions=record.interactions)
            em.SymmetrizeInteractions(log=None)
            p = model.CalculateProbabilitiesBatch(grid_all)
            if world > 1:
                t = torch.from_numpy(p).to(device)
                dist.all_reduce(t)
                p = (t / world).cpu().numpy()
            return p

        units_per_step = float(len(grid_all)) * chains * (nequil + nprod) * nmoves
        metric, unit = "mc_trial_moves_per_s", "moves/s"
        launch_info = plan
        workload = "%s (%d sites / %d instances / %d pairs) MC titration curve pH 0..20 step 0.5 (41 pts), nequil %d, nprod %d, %d chains per pH per GPU" % (
            record.name, record.nsites, record.ninstances, npairs, nequil, nprod, chains)
        h2d = record.ninstances * ((record.ninstances + 3) // 4 * 4) * 8 + record.ninstances * 20 + record.nsites * 8 + len(grid_all) * 8
        d2h = len(grid_all) * (record.ninstances * 8 + 4 * 8 + 8)
        # reference algorithm: nsites single moves cost 20*nsites bytes, npairs double moves 36*nsites
        algo_bytes_per_unit = (record.nsites * 20.0 * record.nsites + npairs * 36.0 * record.nsites) / nmoves

    # ---- value: device-resident ----------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    lib.pce_launch_count(1)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    events = []
    for _ in range(args.steps):
        flush.fill_(1)                                # L2 flush between timed steps (outside the timed events)
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        step_device()
        stop.record(stream)
        events.append((start, stop))
    barrier()
    clock_info = clocks.stop() if rank == 0 else None
    launches = int(lib.pce_launch_count(0))
    elapsed_ms = sum(a.elapsed_time(b) for a, b in events)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_units = units_per_step * args.steps * (world if args.workload != "synthA" else 1)
    value = total_units / (elapsed_ms * 1e-3)

    # ---- e2e: public API from host buffers -----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        for _ in range(min(args.warmup, 2)):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": total_units / float(t.item()), "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as handle:
                peaks = json.load(handle)
        except (OSError, ValueError):
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        kernel_ms = elapsed_ms / max(args.steps, 1)
        # bytes this algorithm actually moves at HBM per launch: the model staged once per CTA wave (L2 serves the
        # repeats) plus the outputs; the working set lives in shared memory (fixtures) or L2 (synthM)
        actual_bytes = float(h2d + d2h)
        on_chip = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_summary.json")) as handle:
                summary = json.load(handle)
            key = "k_enum" if args.workload == "synthA" else ("k_mc_warp" if launch_info.get("kernel") == 2 else "k_mc_thread")
            entry = key + "_ifp20" if key == "k_mc_warp" and args.workload != "synthM" else key
            on_chip = dict(summary[entry], kernel=key)
        except (OSError, ValueError, KeyError):
            pass
        roofline = {
            "bound": "hbm", "achieved": actual_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": actual_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "traffic": (on_chip or {}).get("dram_bytes_per_launch"),
            "on_chip": on_chip,
            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "note": "HBM is not the limiter of this path (working set in shared memory / L2): `achieved` counts the bytes this algorithm "
                    "actually moves at HBM per launch, so `frac` is ~0 by design; the real limiter and its ncu-measured utilisation are in "
                    "`on_chip`.  reference_algorithm_effective_gbs = the reference algorithm's gather traffic (SURVEY.md 8d) replaced by this run.",
            # the binding on-chip resource and its measured utilisation (ncu, profiles/): issue slots for the chain-per-thread
            # and enumeration kernels, the L1TEX/LSU wavefront pipe for the chain-per-warp kernel on large models
            "limiter": (on_chip or {}).get("limiter"),
            "limiter_frac": (on_chip or {}).get("l1tex_throughput_frac" if launch_info.get("kernel") == 2 and args.workload == "synthM" else "issue_active_frac"),
            "reference_algorithm_bytes_per_unit": algo_bytes_per_unit,
            "reference_algorithm_effective_gbs": algo_bytes_per_unit * value / world / 1e9,
            "kernel_ms_per_launch": kernel_ms,
        }
        if args.workload != "synthA" and launch_info.get("kernel") == 2 and on_chip is not None:
            # chain-per-warp kernel: the difference rows are served by L2.  Algorithmic L2 bytes of one launch = one
            # difference row per changed site of every accepted move + the four W entries of every double-move proposal
            # + the rows read by the field initialisation; accepted moves are counted in production scans only and
            # scaled to all scans.  The denominator is measured here with a streaming read of an L2-resident buffer.
            c = d_counters.sum(dim=0).cpu().numpy().astype(float)
            scale = float(nequil + nprod) / float(nprod)
            row_bytes = (record.ninstances - record.nsites) * 8.0          # fields relative to the first instance of each site
            l2_bytes = scale * ((c[1] + 2.0 * c[3]) * row_bytes + c[2] * 32.0) + float(len(grid_all)) * chains * record.nsites * row_bytes
            peak = C.c_double(0.0)
            _native.check(lib.pce_probe_read_bandwidth(local_rank, 48 * 1024 * 1024, 40, C.byref(peak)))
            on_chip["l2_algorithmic_bytes_per_launch"] = l2_bytes
            on_chip["l2_algorithmic_gbs"] = l2_bytes / (kernel_ms * 1e-3) / 1e9
            on_chip["l2_read_peak_gbs_measured"] = peak.value
            on_chip["l2_frac"] = on_chip["l2_algorithmic_gbs"] / max(peak.value, 1e-9)
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak" if args.workload != "synthA" else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "reference fixture model, synthetic chains" if args.workload in ("lysozyme", "ifp20", "defensin") else "synthetic",
            "config": {"workload": workload, "l2": "flushed between timed steps (256 MiB write)", "parallelism": "chains sharded over %d GPU(s), one all-reduce of counts" % world,
                       "launch": launch_info},
            "clocks": clock_info, "gpu_launches": launches, "roofline": roofline,
        }
        if e2e is not None:
            line["e2e"] = e2e
        if args.workload != "synthA":
            c = d_counters.sum(dim=0).cpu().numpy().astype(float)       # production trials: moves, accepted, flips, accepted
            line["config"]["acceptance"] = {"single": c[1] / max(c[0], 1.0), "double": c[3] / max(c[2], 1.0),
                                            "overall": (c[1] + c[3]) / max(c[0] + c[2], 1.0)}
        if world == 1 and not args.no_e2e and args.workload in ("lysozyme", "ifp20", "defensin"):
            # BASELINE's third metric: wall time of ONE titration curve through the public API (TitrationCurves), host
            # arrays in, curves out.  (a) the reference's statistical effort per pH value (20000 production scans) split
            # over 64 chains; (b) the package defaults (64 chains x 20000 scans = 64x the reference's samples).
            curve = {}
            for label, nch, npr in (("reference_effort", 64, 313), ("defaults", 64, 20000 if args.workload != "ifp20" else 2000)):
                cm = P.MEADModel(record, log=None, device=local_rank)
                cm.DefineMCModel(P.MCModelDefault(nequil=500, nprod=npr, randomSeed=7, nchains=nch), log=None)
                P.TitrationCurves(cm, log=None, curveSampling=0.5, curveStart=0.0, curveStop=20.0).CalculateCurves(log=None)     # warm-up
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tc = P.TitrationCurves(cm, log=None, curveSampling=0.5, curveStart=0.0, curveStop=20.0)
                tc.CalculateCurves(log=None)
                tc.CalculateHalfpKs()
                curve[label] = {"wall_ms": 1000.0 * (time.perf_counter() - t0), "chains": nch, "nequil": 500, "nprod_per_chain": npr, "ph_points": tc.nsteps}
            line["titration_curve"] = curve
        if world == 1 and not args.no_cpu_baseline and args.workload == "synthA":
            # the reference cannot run this workload at all; its per-state cost is nsites (nsites - 1) / 2 gathers + one exp per
            # pH value, so the 24-site rate is an UPPER bound of what it would do on 32 sites (276 against 496 gathers)
            rate, seconds, kind, nstates = cpu_baseline_analytic(24, 7.0)
            line["cpu_baseline"] = {"value": rate, "unit": "states/s", "cores": 1, "kind": kind,
                                    "sample": "24-site variant of the same generator, 2^24 states at ONE pH value, %.1f s on one host core; the "
                                              "reference caps exhaustive enumeration at 2^26 states and needs one pass per pH value "
                                              "(x 41 for this curve), so 2^32 x 41 pH extrapolates to >= %.0f core-hours" % (
                                                  seconds, 2.0 ** 32 * 41 / rate / 3600.0)}
        if world == 1 and not args.no_cpu_baseline and args.workload == "synthM":
            # 20 + 100 scans of one chain at pH 7: ~8e5 trial moves, 5-10 s on one host core
            rate, seconds, kind = cpu_baseline(record, 20, 100, grid_all[14:15], 1)
            line["cpu_baseline"] = {"value": rate, "unit": "moves/s", "cores": 1, "kind": kind,
                                    "sample": "one single-chain run at pH 7 (20 + 100 scans of %d trial moves, MT19937), %.1f s on one host core" % (
                                        nmoves, seconds)}
        if world == 1 and not args.no_cpu_baseline and args.workload in ("lysozyme", "ifp20", "defensin"):
            ncurves = 6 if args.workload != "ifp20" else 1            # 10-20 s of one host core
            rate, seconds, kind = cpu_baseline(record, 500, 20000, grid_all, 1, repeats=ncurves)
            line["cpu_baseline"] = {"value": rate, "unit": "moves/s", "cores": 1, "kind": kind,
                                    "sample": "%d single-chain %s curve(s) (41 pH, 500+20000 scans, MT19937), %.1f s on one host core" % (
                                        ncurves, record.name, seconds)}
            if "titration_curve" in line:
                line["titration_curve"]["reference_one_core_wall_ms"] = 1000.0 * seconds / ncurves
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
