# SPDX-License-Identifier: MIT
"""torchrun smoke test of the library's NCCL binding (mk_comm_init_rank / mk_allreduce_stats):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/comm_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import markov_b200 as mm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a): print(f"[rank {rank}]", *a, file=sys.stderr, flush=True)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
say("process group up; MK_NCCL_LIB =", os.environ.get("MK_NCCL_LIB"))
comm = mm.sharding.Communicator(rank, world, local)
say("communicator up")
x = torch.arange(10, dtype=torch.float64, device="cuda") * (rank + 1)
comm.allreduce_(x)
torch.cuda.synchronize()
want = torch.arange(10, dtype=torch.float64, device="cuda") * (world * (world + 1) / 2)
assert torch.equal(x, want), (x, want)
say("all-reduce ok")
comm.close()
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("comm_check ok", world)
