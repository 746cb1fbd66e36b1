"""Pieces of the data-parallel step with the PUSH exchange (DataParallelTower._step_staged_push), CUDA events between them
on every rank, max over ranks, eager issue (compute half replayed as the C-side graph).
python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 profiles/dp_push_timeline.py [C2]"""
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
t = DSSMTower(conf, max_nnz=b.nnz + 1024, params=init_params(conf, 0), symmetric=True)
dp = DataParallelTower(t, comm="nvlink")
assert dp.exchange == "push"
t.stage(t.to_device(b))
t.capture_graph_dp()
main = torch.cuda.current_stream()
own = ptr(dp._flag_buf)
c, n = conf, world
names = ["fwd_bwd", "small_allreduce_issue", "gather_push", "flags_0", "owner_pass", "small_grads_wait+adam", "flags_1"]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
acc = [0.0] * len(names)
N, W = 30, 5
for it in range(N + W):
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    t.fwd_bwd_begin_staged(); ev[1].record()
    w_rest = dist.all_reduce(t.comm[dp.w1_end:], op=dist.ReduceOp.AVG, async_op=True); ev[2].record()
    check(lib.dssm_tower_backward_w1_push(t._h, dp._peer_slots, dp._peer_valid, dp._epoch_ptr, n, rank, dp._per, stream_ptr(main))); ev[3].record()
    check(lib.dssm_peer_signal(dp._peer_flags, n, rank, 0, 2, stream_ptr(main)))
    check(lib.dssm_peer_wait(own, n, 0, 2, stream_ptr(main))); ev[4].record()
    check(lib.dssm_w1_slots_reduce_adam(ptr(dp._slots), ptr(dp._valid), dp._epoch_ptr, dp._peer_w, dp._mc_w if dp.use_multicast else None, n, rank,
                                        c.TRIGRAM_D, c.layers[0], dp._per, ptr(t.m), ptr(t.v), ptr(t.beta_pow), c.learning_rate, c.beta1, c.beta2,
                                        c.adam_eps, stream_ptr(main))); ev[5].record()
    w_rest.wait()
    t.adam_range(dp.w1_end, t.P - dp.w1_end, 1.0)
    t.adam_advance(); ev[6].record()
    check(lib.dssm_peer_signal(dp._peer_flags, n, rank, 1, 2, stream_ptr(main)))
    check(lib.dssm_peer_wait(own, n, 1, 2, stream_ptr(main)))
    check(lib.dssm_peer_epoch_advance(own, stream_ptr(main))); ev[7].record()
    torch.cuda.synchronize()
    if it >= W:
        for i in range(len(names)):
            acc[i] += ev[i].elapsed_time(ev[i + 1]) * 1e3 / N
tt = torch.tensor(acc, device="cuda", dtype=torch.float64)
mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
mn = tt.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"{name} push exchange, n={world}, multicast={dp.use_multicast}: us per piece, max over ranks (min)")
    for nm, a, b_ in zip(names, mx.tolist(), mn.tolist()):
        print(f"  {nm:26s} {a:8.1f}  ({b_:.1f})")
    print(f"  sum {sum(mx.tolist()):.1f}")
dist.destroy_process_group()
