# SPDX-License-Identifier: MIT
"""Where does a multi-rank step lose time?  cfg 3 per rank; step variants timed back to back (torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import markov_b200 as mm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
K = mm.LogSemiring[np.float32]
B, T, D = 128, 150, 3000
fsm, pdf = mm.graphs.denominator(K)
c = mm.compile(fsm, mm.statemap(fsm, D, pdf)); b = mm.batch(*[c] * B)
V = (torch.randn((B, T, D), generator=torch.Generator(device="cuda").manual_seed(303 + rank), device="cuda") * 2).permute(0, 2, 1)
post = torch.empty((T, D, B), device="cuda"); ttl = torch.empty((B,), device="cuda")
stats = torch.zeros(D + 2, dtype=torch.float64, device="cuda")
comm = mm.sharding.Communicator(rank, world, local)
def v_none(): mm.pdfposteriors(b, V, out=(post, ttl), stats=stats)
def v_mk(): mm.pdfposteriors(b, V, out=(post, ttl), stats=stats); comm.allreduce_(stats)
def v_torch(): mm.pdfposteriors(b, V, out=(post, ttl), stats=stats); dist.all_reduce(stats)
def v_nostats(): mm.pdfposteriors(b, V, out=(post, ttl))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
modes = [("no exchange", v_none, 0, 0), ("mk_allreduce_stats", v_mk, 0, 0), ("torch all_reduce", v_torch, 0, 0),
         ("mk + kernel events", v_mk, 1, 0), ("mk + NVML sampler", v_mk, 0, 1), ("mk + both", v_mk, 1, 1), ("mk_allreduce_stats", v_mk, 0, 0)]
for name, fn, prof, samp in modes:
    b.profile(bool(prof))
    sampler = bench.ClockSampler(local) if (samp and rank == 0) else None
    if sampler: sampler.start()
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(10): fn()
    t_enq = time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if sampler: sampler.result()
    if rank == 0: print(f"{name:22s} {float(ms):7.3f} ms/step (max over ranks), host enqueue {1e3 * t_enq / 10:6.3f} ms/step", flush=True)
dist.barrier(); dist.destroy_process_group()
