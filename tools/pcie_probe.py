# SPDX-License-Identifier: MIT
"""Host <-> device copy bandwidth per rank, alone and with every rank copying at once (torchrun): what bounds the
host-buffer (e2e) leg of bench.py at N > 1.  230 MB pinned buffers, as one bench step moves each way."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
binding = bench.bind_rank_to_cores(torch, local, int(os.environ.get("LOCAL_WORLD_SIZE", world))) if os.environ.get("BIND", "1") == "1" else "unbound"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 230_400_000 // 4
h_in, h_out = torch.empty(n, pin_memory=True), torch.empty(n, pin_memory=True)
d_in, d_out = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
h_in.normal_()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(mode, reps=8):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if mode in ("h2d", "both"):
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if mode in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return reps * n * 4 / (time.perf_counter() - t0) / 1e9
for mode in ("h2d", "d2h", "both"):
    run(mode, 2)
    dist.barrier()
    alone = run(mode) if rank == 0 else 0.0      # rank 0 alone
    dist.barrier()
    together = run(mode)                          # every rank at once
    t = torch.tensor([alone, together], device="cuda")
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        tg = [float(x[1]) for x in allt]
        print(f"{mode:5s} per direction: rank 0 alone {float(allt[0][0]):6.1f} GB/s; all {world} ranks at once: min {min(tg):6.1f} mean {sum(tg) / world:6.1f} max {max(tg):6.1f} GB/s per rank ({binding})", flush=True)
dist.barrier(); dist.destroy_process_group()
