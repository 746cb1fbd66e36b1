"""Where the data-parallel step (comm='nvlink') spends its time: CUDA events between the pieces of
DataParallelTower._step_staged_nvlink, eager (compute half as the C graph).  Launch with torchrun, 2+ ranks:
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 profiles/dp_timeline.py [C2]"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from dssm_b200 import DSSMTower, baseline_config
from dssm_b200._lib import check, lib, ptr, stream_ptr
from dssm_b200.parallel import DataParallelTower
from dssm_b200.synthetic import init_params, make_batch

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
conf = baseline_config(name)
b = make_batch(conf, seed=rank)
t = DSSMTower(conf, max_nnz=b.nnz, params=init_params(conf, 0), symmetric=True)
dp = DataParallelTower(t, comm="nvlink")
t.stage(t.to_device(b))
t.capture_graph_dp()
c = conf
labels = ["fwd+bwd (C graph)", "small all-reduce issue + dW1 gather", "barrier 1", "shard pull/Adam/push", "wait small all-reduce",
          "Adam small + advance", "barrier 2"]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(labels) + 1)]

def step():
    ev[0].record()
    t.fwd_bwd_begin_staged(); ev[1].record()
    w = dist.all_reduce(t.comm[dp.w1_end:], op=dist.ReduceOp.AVG, async_op=True)
    t.backward_w1(0, 1); ev[2].record()
    dp._h_comm.barrier(channel=0); ev[3].record()
    check(lib.dssm_w1_shard_reduce_adam(dp._peer_dw, dp._peer_w, world, rank, c.TRIGRAM_D, c.layers[0], dp.row_begin, dp.row_end,
                                        ptr(t.m), ptr(t.v), ptr(t.beta_pow), c.learning_rate, c.beta1, c.beta2, c.adam_eps, stream_ptr()))
    ev[4].record()
    w.wait(); ev[5].record()
    t.adam_range(dp.w1_end, t.P - dp.w1_end, 1.0); t.adam_advance(); ev[6].record()
    dp._h_params.barrier(channel=0); ev[7].record()

acc = [0.0] * len(labels)
N = 20
for i in range(N + 5):
    step()
    torch.cuda.synchronize()
    if i >= 5:
        for j in range(len(labels)):
            acc[j] += ev[j].elapsed_time(ev[j + 1]) * 1e3 / N
if rank == 0:
    print(f"{name} N={world} data-parallel step, us: " + "  ".join(f"{l}={a:.1f}" for l, a in zip(labels, acc)) + f"  | sum {sum(acc):.1f}", flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
