import os, torch, torch.distributed as dist
r=int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
for mb in (3.75, 7.5, 15, 30, 59.6):
    n=int(mb*1e6/4); x=torch.ones(n,device="cuda")
    for _ in range(5): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0,e1=torch.cuda.Event(True),torch.cuda.Event(True)
    e0.record()
    for _ in range(20): dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    if r==0: print(f"allreduce {mb} MB: {e0.elapsed_time(e1)/20*1e3:.1f} us")
dist.destroy_process_group()
