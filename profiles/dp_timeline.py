"""Where the data-parallel step (comm='nvlink', chunked exchange) spends its time: CUDA events around the pieces of
DataParallelTower._step_staged_nvlink, eager (compute half as the C graph), in three arrangements:
  serial    gather chunk k, signal, wait, owner pass of chunk k, all on one stream (no overlap)
  gather    the K gather chunks alone, back to back
  pipelined what the step does: owner pass of chunk k on the exchange stream under the gather of chunk k+1
Launch with torchrun, 2+ ranks:
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 profiles/dp_timeline.py [C2] [K]"""
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
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
os.environ["DSSM_DP_CHUNKS"] = str(K)
conf = baseline_config(name)
b = make_batch(conf, seed=rank)
t = DSSMTower(conf, max_nnz=b.nnz, params=init_params(conf, 0), symmetric=True)
dp = DataParallelTower(t, comm="nvlink")
t.stage(t.to_device(b))
t.capture_graph_dp()
main = torch.cuda.current_stream()
own = ptr(dp._flag_buf)
stride = K + 1


def timed(fn, n=20, warm=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for i in range(n + warm):
        t.fwd_bwd_begin_staged()
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            tot += e0.elapsed_time(e1) * 1e3 / n
    return tot


def finish():
    check(lib.dssm_peer_signal(dp._peer_flags, world, rank, K, stride, stream_ptr(main)))
    check(lib.dssm_peer_wait(own, world, K, stride, stream_ptr(main)))
    check(lib.dssm_peer_epoch_advance(own, stream_ptr(main)))


def gather_only():
    for k in range(K):
        t.backward_w1(k, K)


def serial():
    for k in range(K):
        t.backward_w1(k, K)
        check(lib.dssm_peer_signal(dp._peer_flags, world, rank, k, stride, stream_ptr(main)))
        check(lib.dssm_peer_wait(own, world, k, stride, stream_ptr(main)))
        dp._exchange_rows(*dp._owned[k])
    finish()


def exchange_only():  # dW1 as the last gather left it
    for k in range(K):
        check(lib.dssm_peer_signal(dp._peer_flags, world, rank, k, stride, stream_ptr(main)))
        check(lib.dssm_peer_wait(own, world, k, stride, stream_ptr(main)))
        dp._exchange_rows(*dp._owned[k])
    finish()


def pipelined():
    dp._xstream.wait_stream(main)
    for k in range(K):
        t.backward_w1(k, K)
        check(lib.dssm_peer_signal(dp._peer_flags, world, rank, k, stride, stream_ptr(main)))
        with torch.cuda.stream(dp._xstream):
            check(lib.dssm_peer_wait(own, world, k, stride, stream_ptr(dp._xstream)))
            dp._exchange_rows(*dp._owned[k])
    main.wait_stream(dp._xstream)
    finish()


def flags_only():
    for k in range(K):
        check(lib.dssm_peer_signal(dp._peer_flags, world, rank, k, stride, stream_ptr(main)))
        check(lib.dssm_peer_wait(own, world, k, stride, stream_ptr(main)))
    finish()


res = {}
for nm, fn in (("gather chunks only", gather_only), ("flags only (K+1 signal/wait pairs)", flags_only), ("exchange only", exchange_only),
               ("serial gather+exchange", serial), ("pipelined", pipelined)):
    res[nm] = timed(fn)
if rank == 0:
    print(f"{name} N={world} K={K} multicast={dp.use_multicast}, us after fwd+bwd: " + "  ".join(f"{k}={v:.1f}" for k, v in res.items()), flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
