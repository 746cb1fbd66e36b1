"""Isolates what bounds dssm_w1_shard_reduce_adam: the same launch with (1) every pointer local, (2) remote loads
only, (3) remote stores only, (4) the real tables.  torchrun, >= 2 ranks."""
import ctypes as C, os, sys
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
sys.path.insert(0, ".")
from dssm_b200._lib import check, lib, ptr, stream_ptr

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
D, L1 = 49284, 300
n = D * L1
W = symm_mem.empty(n, dtype=torch.float32, device="cuda").normal_()
G = symm_mem.empty(n, dtype=torch.float32, device="cuda").normal_()
hW, hG = symm_mem.rendezvous(W, dist.group.WORLD), symm_mem.rendezvous(G, dist.group.WORLD)
m, v = torch.zeros(n, device="cuda"), torch.ones(n, device="cuda")
bp = torch.tensor([0.9, 0.999], device="cuda")
per = (D + world - 1) // world
r0, r1 = rank * per, min(D, (rank + 1) * per)
arr = C.c_void_p * world
real_w, real_g = [int(p) for p in hW.buffer_ptrs], [int(p) for p in hG.buffer_ptrs]
self_w, self_g = [real_w[rank]] * world, [real_g[rank]] * world
cases = {"all local": (self_g, self_w), "remote loads only": (real_g, self_w), "remote stores only": (self_g, real_w), "real": (real_g, real_w)}
for name, (pg, pw) in cases.items():
    tg, tw = arr(*pg), arr(*pw)
    def run():
        check(lib.dssm_w1_shard_reduce_adam(tg, tw, world, rank, D, L1, r0, r1, ptr(m), ptr(v), ptr(bp), 0.01, 0.9, 0.999, 1e-8, stream_ptr()))
    for _ in range(3): run()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"N={world} shard rows {r1 - r0}: {name:20s} {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us", flush=True)
    dist.barrier()
torch.cuda.synchronize()
os._exit(0)
