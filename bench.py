# SPDX-License-Identifier: MIT
"""bench.py — LF-MMI denominator forward-backward throughput (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm
                                                             # (oracle port) on the host cores

A step is one `pdfposteriors` (forward + backward + pdf posteriors, src/inference.jl:145-161)
over one batch of synthetic input.  Workload per GPU = BASELINE.json configs[2]: the synthetic
Kaldi-chain denominator graph (S = 30 000 states, ~520k arcs, D = 3 000 pdfs; SURVEY.md §8d cfg 3),
B = 128 utterances x T = 150 frames, LogSemiring{Float32}.  With N GPUs every rank runs its own
128 utterances on a replicated graph (weak scaling; N = 8 is configs[3]'s 1 024 utterances) and
the ranks exchange only [Σ logZ, #frames, pdf occupancy[D]] with one NCCL all-reduce per step.

Prints ONE JSON line (rank 0).  `value`: device-resident inputs; `e2e`: the same call on HOST
(pinned) buffers, H2D + D2H inside the timed region; `roofline`: the shared-graph kernel's
algorithmic HBM bytes / its CUDA-event time against MEASURED_PEAKS.json; `cpu_baseline`: the
oracle port timed on this box's cores on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LF-MMI denominator forward-backward frames/sec (batch x T)"
UNIT = "frames/s"
B_PER_GPU, T_FRAMES, N_PDF, N_TOKENS, SEED = 128, 150, 3000, 15000, 303
HBM_FALLBACK_GBS = 6650.0
# dram__bytes_read.sum + dram__bytes_write.sum of the two shared_fb_kernel launches of ONE pdfposteriors call on
# this workload, from the ncu capture summarised in profiles/r02_shared_fb_kernel_ncu.md (round-2 end-of-round build:
# forward 0.219 + 2.289 GB, backward 2.767 + 1.679 GB); null for any other shape
NCU_DRAM_BYTES_PER_LAUNCH = 6.954e9


def workload_config(n_gpus, b_per_gpu, frames):
    return {"workload": "BASELINE.json configs[2]: synthetic Kaldi-chain denominator graph (30000 states, "
                        "3000 pdfs, SURVEY.md 8d cfg 3), LogSemiring{Float32}, forward-backward (pdfposteriors)",
            "batch_per_gpu": b_per_gpu, "global_batch": b_per_gpu * n_gpus, "frames": frames,
            "graph": "replicated per GPU", "parallelism": f"utterance-sharded x{n_gpus}",
            "l2": "inputs larger than L2 (alpha store 2.3 GB, emissions 230 MB per step)"}


class ClockSampler(threading.Thread):
    """NVML: SM clock + throttle reasons every few milliseconds while the timed region runs.  NVML is
    initialised (and queried once: the first query is slow) by the constructor, before the timed region; the
    thread only polls."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.mask, self.max_mhz, self.err, self.h, self.stop_flag = [], 0, None, None, None, False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self._query()  # warm the query path
            self.samples, self.mask = [], 0
        except Exception as e:  # no NVML: report it rather than invent numbers
            self.err = repr(e)

    def _query(self):
        self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
        try:
            self.mask |= self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            self.mask |= self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)

    def run(self):
        if self.h is None:
            return
        try:
            while not self.stop_flag:
                self._query()
                time.sleep(0.003)
        except Exception as e:
            self.err = repr(e)

    def result(self):
        self.stop_flag = True
        if self.is_alive():
            self.join(timeout=2)
        if self.err or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "error": self.err or "no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": [n for bit, n in self.REASONS.items() if self.mask & bit], "samples": len(self.samples)}


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback"


def build_graph(mm):
    K = mm.LogSemiring[np.float32]
    return K, mm.graphs.denominator(K, n_tokens=N_TOKENS, n_pdf=N_PDF, seed=SEED)


def cpu_sample(fsm, pdfids, n_utts, frames, threads, seed=SEED):
    """Time the oracle port on `n_utts` utterances x `frames` frames of the workload."""
    import oracle
    og = oracle.OracleGraph(fsm, pdfids, N_PDF)
    rng = np.random.default_rng(seed)
    V = (rng.standard_normal((n_utts, frames, N_PDF)) * 2).astype(np.float32)
    t0 = time.perf_counter()
    oracle.pdfposteriors([og] * n_utts, V, threads=threads)
    return n_utts * frames / (time.perf_counter() - t0)


def run_reference(args):
    """The reference's own CPU algorithm for the path (oracle port: the reference is Julia and
    cannot run here) on the host cores; rank 0 only.  Two figures: ALL cores this process may use (one
    full-length utterance per core and step — torchrun's OMP_NUM_THREADS=1 is ignored on purpose) = the
    line's `value`, and ONE core (faithful to the reference, whose src/ has no threads)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import markov_b200 as mm
    import oracle
    K, (fsm, pdfids) = build_graph(mm)
    threads = oracle.num_threads()
    frames = T_FRAMES  # full 150-frame utterances (about 7 s each on one core: a step is one utterance per core)
    og = oracle.OracleGraph(fsm, pdfids, N_PDF)
    rng = np.random.default_rng(SEED)
    V = (rng.standard_normal((threads, frames, N_PDF)) * 2).astype(np.float32)
    t0 = time.perf_counter()
    oracle.pdfposteriors([og], V[:1], threads=1)
    single = frames / (time.perf_counter() - t0)
    # bounded: the whole run (warm-up included) stays within a few minutes whatever --steps says
    est_step = frames / single * 1.3
    steps = max(1, min(args.steps, int(150.0 / est_step) - 1))
    warmup = 1 if args.warmup > 0 else 0
    for _ in range(warmup):
        oracle.pdfposteriors([og] * threads, V, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle.pdfposteriors([og] * threads, V, threads=threads)
    dt = time.perf_counter() - t0
    value = steps * threads * frames / dt
    sample = (f"{threads} utterances x {frames} frames of the workload per step, one utterance per core, {steps} timed "
              f"step(s) (bounded from --steps {args.steps})")
    cfg = workload_config(args.gpus, B_PER_GPU, T_FRAMES)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "single_thread": {"value": single, "unit": UNIT, "cores": 1,
                                           "sample": f"1 utterance x {frames} frames"},
                         "build": oracle.build_flags()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle port (C++ restatement of src/inference.jl, OpenMP over utterances); the Julia "
                "reference cannot run in this image"}))


def bind_rank_to_cores(torch, local, n_local):
    """One core set per rank, on the NUMA node of the rank's GPU when sysfs says which: the ranks' launch threads
    stop competing for the same cores, and the pinned host buffers of the e2e leg (first touched after this) land in
    memory local to the GPU's PCIe root.  Returns a short description for the JSON line."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return "unbound"
    near = []
    try:
        pr = torch.cuda.get_device_properties(local)
        path = f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0/local_cpulist"
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            near.extend(range(int(lo), int(hi or lo) + 1))
        near = [c for c in near if c in allowed]
    except Exception:
        near = []
    pool = near if len(near) >= n_local else allowed
    # the ranks that share this pool take disjoint slices of it
    per = max(1, len(pool) // n_local)
    mine = pool[(local % n_local) * per:(local % n_local) * per + per] or pool
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return "unbound"
    return f"{len(mine)} cores ({'GPU-local NUMA node' if near and pool is near else 'even split'})"


def run_ours(args):
    import torch
    import markov_b200 as mm
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (libmarkov_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    binding = "unbound"
    try:
        affinity0 = os.sched_getaffinity(0)
    except AttributeError:
        affinity0 = None
    # (also at N = 1: where the pinned buffers of the e2e leg land decides its PCIe rate — the same call measured
    # 7.8 and 8.5 ms per step on two runs of an unbound process; the CPU baseline gets the full core set back below)
    if os.environ.get("MK_BENCH_BIND", "1") != "0":
        binding = bind_rank_to_cores(torch, local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, T, D = args.batch, args.frames, N_PDF
    K, (fsm, pdfids) = build_graph(mm)
    cfsm = mm.compile(fsm, mm.statemap(fsm, D, pdfids))
    bfsm = mm.batch(*[cfsm] * B)
    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    V = torch.randn((B, T, D), generator=gen, device="cuda") * 2  # the network's (B, T, D) output
    Vd = V.permute(0, 2, 1)                                        # (B, D, T) view the API takes
    post = torch.empty((T, D, B), device="cuda")
    ttl = torch.empty((B,), device="cuda")
    lib = mm.lib()
    last = {}

    # the data-parallel exchange: [Σ logZ, #frames, pdf occupancy[D]] (~24 KB) written by the library itself
    # (mk_pdfposteriors_stats) and summed over the ranks by its own NCCL binding (mk_allreduce_stats) on the same
    # stream: no eager reduction, no host synchronisation inside a step
    stats = torch.zeros(D + 2, dtype=torch.float64, device="cuda")
    comm = mm.sharding.Communicator(rank, world, local) if world > 1 else None

    def step():
        mm.pdfposteriors(bfsm, Vd, out=(post, ttl), stats=stats)
        if comm is not None:
            comm.allreduce_(stats)
        last["stats"] = stats

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    # Everything slow and rank-specific happens BEFORE the barrier that opens the timed region: NVML initialisation on
    # rank 0 takes tens of milliseconds, and a rank that enters the region late makes every other rank wait for it in
    # the first all-reduce (round 1 measured exactly that: +1.9 ms per step at N = 8, none of it in the exchange).
    bfsm.profile(True)
    sampler = ClockSampler(local) if rank == 0 else None
    lib.mk_launch_count(1)
    if sampler:
        sampler.start()  # (NVML is already initialised: the first sample lands within microseconds)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()  # the steps were enqueued asynchronously: the GPU is inside the timed region now
    sync_all()
    launches = int(lib.mk_launch_count(0))
    clocks = sampler.result() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    kern = bfsm.kernel_ms(min(args.steps, 64))
    kms = torch.tensor([float(np.mean(kern))], device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_total, kernel_ms = float(ms), float(kms)
    bfsm.profile(False)
    logz_mean = float(last["stats"][0]) / (B * world)

    # ---- e2e: the public API on HOST buffers, copies inside the timed region
    Vh = torch.empty((B, T, D), pin_memory=True)
    Vh.copy_(V)
    post_h = torch.empty((T, D, B), pin_memory=True)
    ttl_h = torch.empty((B,), pin_memory=True)
    Vh_np, post_np, ttl_np = Vh.numpy().transpose(0, 2, 1), post_h.numpy(), ttl_h.numpy()

    def step_e2e():
        mm.pdfposteriors(bfsm, Vh_np, out=(post_np, ttl_np))  # H2D + kernels + D2H + sync inside
        return float(ttl_np.sum())

    for _ in range(2):
        step_e2e()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    single_t = torch.tensor([time.perf_counter() - t0], device="cuda")
    # The same steps with three calls in flight: three batch objects (one graph) take turns, so that step k+1's and
    # k+2's host-to-device copies overlap step k's sweeps and step k-1's device-to-host copy (PCIe is full duplex; with
    # two objects the copies of one batch — 2 x 4.2 ms — no longer fit beside the other's 7.4 ms of sweeps).  Every step
    # still copies its inputs in from pinned memory and its posteriors and log-likelihoods out, and the host reads
    # each result.
    DEPTH = 3
    bf2 = [bfsm] + [mm.batch(*[cfsm] * B) for _ in range(DEPTH - 1)]
    outs = [(post_np, ttl_np)]
    keep_pinned = []
    for _ in range(DEPTH - 1):
        ph, th = torch.empty((T, D, B), pin_memory=True), torch.empty((B,), pin_memory=True)
        keep_pinned.append((ph, th))
        outs.append((ph.numpy(), th.numpy()))
    checks = []

    def run_pipelined(n):
        for k in range(n):
            j = k % DEPTH
            if k >= DEPTH:
                bf2[j].wait()
                checks.append(float(outs[j][1].sum()))
            mm.pdfposteriors(bf2[j], Vh_np, out=outs[j], wait=False)
        for k in range(max(0, n - DEPTH), n):
            bf2[k % DEPTH].wait()
            checks.append(float(outs[k % DEPTH][1].sum()))

    run_pipelined(2 * DEPTH)
    sync_all()
    t0 = time.perf_counter()
    run_pipelined(args.steps)
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(single_t, op=dist.ReduceOp.MAX)
    e2e_s, single_s = float(e2e_t), float(single_t)
    torch.testing.assert_close(torch.from_numpy(outs[1][1]).cuda(), ttl, rtol=1e-5, atol=1e-3)
    # two of these batches with DEVICE-resident inputs, one stream each (what overlap buys without PCIe in the way)
    # (SM sharing on: each batch's sweeps run with half the threads per CTA, so the two cooperative kernels are
    # co-resident and one fills the other's grid-barrier waits; it did not pay inside the host pipeline above:
    # 8.39 against 8.24 ms per step when measured with two calls in flight)
    for bb in bf2[:2]:
        bb.set_overlap(True)
    dev_out = [(post, ttl), (torch.empty_like(post), torch.empty_like(ttl))]
    side = [torch.cuda.Stream(), torch.cuda.Stream()]

    def run_two_streams(n):
        for k in range(n):
            with torch.cuda.stream(side[k & 1]):
                mm.pdfposteriors(bf2[k & 1], Vd, out=dev_out[k & 1])

    torch.cuda.synchronize()
    run_two_streams(4)
    sync_all()
    t0 = time.perf_counter()
    run_two_streams(args.steps)
    torch.cuda.synchronize()
    two_t = torch.tensor([time.perf_counter() - t0], device="cuda")
    if dist is not None:
        dist.all_reduce(two_t, op=dist.ReduceOp.MAX)
    two_s = float(two_t)
    for bb in bf2[:2]:
        bb.set_overlap(False)
    h2d = Vh.numel() * 4
    d2h = (post_h.numel() + ttl_h.numel()) * 4
    # the two paths must agree
    torch.testing.assert_close(torch.from_numpy(ttl_np).cuda(), ttl, rtol=1e-5, atol=1e-3)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    frames_total = world * B * T
    value = frames_total * args.steps / (ms_total * 1e-3)
    # roofline of the dominant kernel: algorithmic HBM bytes:
    # w*(2*Ŝ + 3*D̂) per frame.utterance = α written + α read + emissions read twice + posteriors written
    bytes_per_unit = 4 * (2 * fsm.nstates_hat + 3 * (D + 1))
    units = B * T
    achieved = bytes_per_unit * units / (kernel_ms * 1e-3) / 1e9
    peak, which = hbm_peak()
    # the SFU half of the roofline: 2(nnz + S) + S MUFU operations per frame.utterance for the log-domain
    # arithmetic of the reference (one ex2 per arc and one lg2 per row in each sweep, one ex2 per state for gamma),
    # against the ex2 rate measured on this GPU just now
    import ctypes
    sfu_peak = ctypes.c_double(0.0)
    mm._lib.check(lib.mk_measure_sfu_peak(local, ctypes.byref(sfu_peak)))
    nnz_hat = int(fsm.nnz_hat)
    sfu_per_unit = 2 * (nnz_hat + fsm.nstates_hat) + fsm.nstates_hat
    t_hbm = bytes_per_unit * units / (peak * 1e9)
    t_sfu = sfu_per_unit * units / sfu_peak.value
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(world, B, T),
        "e2e": {"value": frames_total * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s / args.steps,
                "how": "host buffers through mk_pdfposteriors_host_begin / mk_batch_wait, three batches in flight: "
                       "every step copies its inputs in and its results out inside the timed region",
                "calls_in_flight": DEPTH,
                "single_call": {"value": frames_total * args.steps / single_s, "unit": UNIT,
                                "ms_per_step": 1e3 * single_s / args.steps,
                                "how": "one blocking mk_pdfposteriors_host call per step (latency of a lone call)"}},
        "two_batches_in_flight": {"value": frames_total * args.steps / two_s, "unit": UNIT,
                                  "ms_per_step": 1e3 * two_s / args.steps,
                                  "how": "device-resident inputs, two batch objects on two streams with mk_batch_set_overlap: "
                                         "their cooperative sweeps are co-resident on every SM (256 threads per CTA each)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (B, T) == (B_PER_GPU, T_FRAMES) else None,
                     "traffic_unit": "bytes per launch pair (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r02_shared_fb_kernel_ncu.md)",
                     "algorithmic_bytes_per_launch": bytes_per_unit * units,
                     "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({which})",
                     "kernel": "shared_fb_kernel<float, Log> (forward sweep + backward sweep: two cooperative "
                               "launches, timed together)", "kernel_ms": kernel_ms,
                     "bytes_per_frame_utt": bytes_per_unit, "units_per_launch": units,
                     "kernel_share_of_step": kernel_ms / (ms_total / args.steps),
                     "sfu": {"ops_per_frame_utt": sfu_per_unit, "algorithmic_ops_per_launch": sfu_per_unit * units,
                             "achieved": sfu_per_unit * units / (kernel_ms * 1e-3) / 1e12, "peak": sfu_peak.value / 1e12,
                             "unit": "Tops/s", "frac": t_sfu / (kernel_ms * 1e-3),
                             "peak_source": "mk_measure_sfu_peak (ex2.approx micro-kernel, measured in this run)",
                             "note": "operations of the reference's log-domain arithmetic; the kernel itself accumulates "
                                     "linear copies with FFMAs and issues MUFU work per state only (DESIGN.md section 4)"},
                     "t_hbm_ms": 1e3 * t_hbm, "t_sfu_ms": 1e3 * t_sfu,
                     "binding_ceiling": "sfu" if t_sfu > t_hbm else "hbm",
                     "frac_of_binding_ceiling": max(t_hbm, t_sfu) / (kernel_ms * 1e-3),
                     "note": "roofline time = max(t_HBM, t_SFU) (BASELINE.md section 3); `frac` is the HBM fraction. The "
                             "kernel is bound by instructions issued per warp and the per-frame grid barrier, not by "
                             "HBM, SFU or the latency of the L2 gathers (DESIGN.md section 4, profiles/)"},
        "mean_logz": logz_mean,
        "host_binding": binding,
    }
    if world == 1 and not args.skip_cpu_baseline:
        import oracle
        if affinity0 is not None:
            try:
                os.sched_setaffinity(0, affinity0)
            except OSError:
                pass
        threads = oracle.num_threads()
        n_utts, frames = threads, T
        v = cpu_sample(fsm, pdfids, n_utts, frames, threads)
        v1 = cpu_sample(fsm, pdfids, 1, frames, 1)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                               "sample": f"{n_utts} utterances x {frames} frames of the workload, one utterance per "
                                         f"thread (OpenMP), C++ oracle port of src/inference.jl",
                               "single_thread": {"value": v1, "unit": UNIT, "cores": 1,
                                                 "sample": f"1 utterance x {frames} frames"},
                               "build": oracle.build_flags()}
    print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_bestpath(args):
    """BASELINE.json configs[4]: TropicalSemiring Viterbi bestpath on the denominator graph, B = 512, T = 500, one GPU
    (a parity-test shape of the headline metric; printed as its own line on request)."""
    import torch
    import markov_b200 as mm
    torch.cuda.set_device(0)
    B, T, D = (512, 500, N_PDF) if args.batch == B_PER_GPU and args.frames == T_FRAMES else (args.batch, args.frames, N_PDF)
    K = mm.TropicalSemiring[np.float32]
    fsm, pdfids = mm.graphs.denominator(K, n_tokens=N_TOKENS, n_pdf=N_PDF, seed=SEED)
    bfsm = mm.batch(*[mm.compile(fsm, mm.statemap(fsm, D, pdfids))] * B)
    V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(505), device="cuda") * 2).permute(0, 2, 1)
    lib = mm.lib()
    for _ in range(max(args.warmup, 3)):
        path, score = mm.bestpath(bfsm, V)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    lib.mk_launch_count(1)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        path, score = mm.bestpath(bfsm, V)   # two calls in flight without a host synchronisation: the call is asynchronous
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    peak, which = hbm_peak()
    # w*D̂ emissions + 4*Ŝ tropical alpha (stored instead of back-pointers) + the path entry
    bytes_per_unit = 4 * (D + 1) + 4 * fsm.nstates_hat + 4
    achieved = bytes_per_unit * B * T / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "TropicalSemiring Viterbi bestpath frames/sec (batch x T)", "value": B * T / (ms * 1e-3), "unit": UNIT,
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[4]: TropicalSemiring Viterbi bestpath on the synthetic denominator "
                               "graph (30000 states, 3000 pdfs)", "batch_per_gpu": B, "frames": T,
                   "l2": "inputs larger than L2 (tropical alpha store %.1f GB)" % (bfsm.workspace_bytes() / 1e9)},
        "gpu_launches": int(lib.mk_launch_count(0)), "clocks": sampler.result(),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "bytes_per_frame_utt": bytes_per_unit, "units_per_launch": B * T,
                     "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({which})",
                     "kernel": "shared_fb_kernel<float, Tropical> forward sweep + backtrace_kernel (whole call timed)"},
        "workspace_bytes": int(bfsm.workspace_bytes()), "mean_score": float(score.mean())}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="utterances per GPU")
    ap.add_argument("--frames", type=int, default=T_FRAMES)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="posteriors", choices=["posteriors", "bestpath"],
                    help="posteriors = the headline metric (default); bestpath = BASELINE.json configs[4], its own line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "bestpath":
        run_bestpath(args)
    else:
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus != world and world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
        run_ours(args)


if __name__ == "__main__":
    main()
