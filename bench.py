#!/usr/bin/env python
"""bench.py -- DSSM train step (fwd + bwd + TF-Adam) throughput in query-groups/s on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference graph (oracle)

One "step" = one pass of the hot path (new_dssm.py:104-217: FC1 SpMM -> BN/FC stack -> Merge_Negative_Doc ->
cosine/softmax/loss -> backward -> Adam over all parameters) over one synthetic batch of query_BS groups per GPU.
Workload: BASELINE.json configs[1] = C2 (TRIGRAM_D=49284, 300-300-128, NEG=4, query_BS=1024 per GPU); N>1 is
data-parallel with the same per-GPU batch (weak scaling) and an NCCL all-reduce of the flat gradient buffer.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train query-groups/s (q+1+NEG docs)"
UNIT = "query-groups/s"


def workload_conf(name: str, gemm_mode: str = "tc_3xtf32"):
    from dssm_b200 import baseline_config

    return baseline_config(name, gemm_mode)


def describe(conf, name, n_gpus, extra=None):
    """The `config` object: identical in both arms (--impl ours / reference) for the same workload and N -- everything
    arm-specific (gemm mode, exchange, graph replay ...) goes into the separate `run` object of our arm."""
    d = {"workload": f"{name}: TRIGRAM_D={conf.TRIGRAM_D}, layers={'-'.join(map(str, conf.layers))}, NEG={conf.NEG}, "
                     f"query_BS={conf.query_BS} per GPU, BN={'on' if conf.use_bn else 'off'}, fwd+bwd+Adam",
         "query_BS_per_gpu": conf.query_BS, "global_query_BS": conf.query_BS * n_gpus, "NEG": conf.NEG,
         "rows_per_step_per_gpu": conf.rows, "parallelism": f"dp{n_gpus}",
         "l2_policy": "inputs larger than L2: every step streams W1+grads+Adam m,v (4 x 59 MB) plus a fresh batch; no explicit flush"}
    if extra:
        d.update(extra)
    return d


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
def threaded_port(ocfg, params, threads: int, dtype=np.float32):
    """The oracle with its three large loops spread over all host threads -- same arithmetic, element for element.
    NumPy ufuncs and SciPy's sparse kernels release the GIL, so X@W1 is split by row blocks, X^T@dh1 by column blocks of
    dh1 and Adam by parameter ranges over a thread pool; the dense layers already use every OpenBLAS thread.  (TensorFlow
    runs MatMul and ApplyAdam on its intra-op pool the same way; its SparseTensorDenseMatMul CPU functor is
    single-threaded, so this port is, if anything, generous to the reference.)"""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import DSSMOracle

    pool = ThreadPoolExecutor(max(threads, 1))

    class ThreadedPort(DSSMOracle):
        def _spmm(self, X, W):
            n = X.shape[0]
            T = max(1, min(threads, n // 128))
            bounds = np.linspace(0, n, T + 1).astype(np.int64)
            out = np.empty((n, W.shape[1]), self.dtype)

            def block(i):
                out[bounds[i]:bounds[i + 1]] = X[bounds[i]:bounds[i + 1]] @ W

            list(pool.map(block, range(T)))
            return out

        def _spmm_t(self, X, dh):
            L = dh.shape[1]
            T = max(1, min(threads, L // 16))
            bounds = np.linspace(0, L, T + 1).astype(np.int64)
            XT = X.T
            out = np.empty((X.shape[1], L), self.dtype)

            def block(i):
                out[:, bounds[i]:bounds[i + 1]] = XT @ np.ascontiguousarray(dh[:, bounds[i]:bounds[i + 1]])

            list(pool.map(block, range(T)))
            return out

        def _adam_tensor(self, k, g, lr_t, b1, b2, eps):
            n = g.size
            if n < (1 << 20) or threads == 1:
                return super()._adam_tensor(k, g, lr_t, b1, b2, eps)
            one = self.dtype.type(1)
            p, m, v = (np.ascontiguousarray(a).reshape(-1) for a in (self.p[k], self.m[k], self.v[k]))
            gf = np.ascontiguousarray(g).reshape(-1)
            bounds = np.linspace(0, n, threads + 1).astype(np.int64)

            def block(i):
                sl = slice(bounds[i], bounds[i + 1])
                m[sl] = b1 * m[sl] + (one - b1) * gf[sl]
                v[sl] = b2 * v[sl] + (one - b2) * (gf[sl] * gf[sl])
                p[sl] = p[sl] - lr_t * m[sl] / (np.sqrt(v[sl]) + eps)

            list(pool.map(block, range(threads)))
            shape = g.shape
            self.p[k], self.m[k], self.v[k] = p.reshape(shape), m.reshape(shape), v.reshape(shape)

    return ThreadedPort(ocfg, params, dtype)


def cpu_port_step_time(conf, batches, params, steps: int, threads: int):
    """The oracle (NumPy/SciPy restatement of new_dssm.py) timed on the host, all threads: `steps` full train steps."""
    from oracle import OracleConfig

    ocfg = OracleConfig(TRIGRAM_D=conf.TRIGRAM_D, layers=tuple(conf.layers), NEG=conf.NEG, query_BS=conf.query_BS,
                        learning_rate=conf.learning_rate, use_bn=conf.use_bn, act=conf.act, loss_eps=conf.loss_eps,
                        loss_div_bs=conf.loss_div_bs)
    orc = threaded_port(ocfg, params, threads)
    Xs = [b.to_scipy() for b in batches]
    orc.train_step(Xs[0])  # warm-up (BLAS thread pools, page faults)
    times = []
    for s in range(steps):
        t0 = time.perf_counter()
        orc.train_step(Xs[s % len(Xs)])
        times.append(time.perf_counter() - t0)
    return times


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: TensorFlow (the reference's engine) is not installable here, so the reference arm is the
    CPU restatement of its graph (oracle/, kind 'port') on all host cores.  Rank 0 only.  Nothing on this arm maps the
    product library (dssm_b200's ctypes binding loads lazily, on the first C call; none is made here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dssm_b200.synthetic import init_params, make_batch

    conf = workload_conf(args.workload)
    batches = [make_batch(conf, seed=s) for s in range(2)]
    params = init_params(conf, 0)
    threads = host_threads()
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is entitled to every host thread
    from threadpoolctl import threadpool_limits

    with threadpool_limits(limits=threads):
        for _ in range(max(args.warmup - 1, 0)):
            cpu_port_step_time(conf, batches[:1], params, 0, threads)
        times = cpu_port_step_time(conf, batches, params, args.steps, threads)
    ms = 1e3 * float(np.mean(times))
    value = conf.query_BS / (ms / 1e3)
    sample = (f"{args.steps} full {args.workload} train steps (query_BS={conf.query_BS}) of the NumPy/SciPy port, one process, sparse "
              f"products / Adam / dense layers on {threads} threads")
    import dssm_b200._lib as _l

    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": describe(conf, args.workload, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "product_library_loaded": _l.loaded()}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
def algorithmic_bytes(conf, nnz, P):
    """SURVEY.md section 8(d) formulas, with the ACTUAL nnz of the generated batches."""
    R, D, L1 = conf.rows, conf.TRIGRAM_D, conf.layers[0]
    widths = list(conf.layers)
    return {
        "spmm_fwd": nnz * (4 * L1 + 8) + R * 4 * L1 + (R + 1) * 4,
        "dw_gather": nnz * (4 * L1 + 8) + D * 4 * L1,
        "adam": 28 * P,
        "dense": R * 4 * sum(widths) * 6,  # c = 6 activation passes
        "cos_loss": 2 * R * 4 * widths[-1],
    }


def dense_flops(conf):
    w = list(conf.layers)
    return 6.0 * conf.rows * sum(a * b for a, b in zip(w[:-1], w[1:]))


def pct(xs, q):
    return float(np.percentile(np.asarray(xs, dtype=np.float64), q))


def epoch_matrices(batches, conf):
    """The three epoch matrices pull_batch slices (utils/utils.py:45-61), built by stacking the batches' parts."""
    import scipy.sparse as sp

    B = conf.query_BS
    Xs = [b.to_scipy() for b in batches]
    q = sp.vstack([X[:B] for X in Xs], format="csr")
    p = sp.vstack([X[B:2 * B] for X in Xs], format="csr")
    n = sp.vstack([X[2 * B:] for X in Xs], format="csr")
    return q, p, n


class Env:
    """Process-wide bench state: rank / world / device and the collective helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, flag: bool) -> bool:
        if self.world == 1:
            return bool(flag)
        t = self.torch.tensor([1 if flag else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


def build_tower(env, conf, batches, params, dp_comm, sync_bn=False):
    """(tower, dp or None).  The choice of symmetric memory is made collectively: every rank falls back together."""
    from dssm_b200 import DSSMTower
    from dssm_b200.parallel import DataParallelTower

    # 5 % head-room: later legs (dp_parity, the loader's epoch) draw fresh batches whose nnz varies by a fraction of a percent
    max_nnz = int(max(b.nnz for b in batches) * 1.05) + 1024
    want_symm = env.world > 1 and dp_comm == "nvlink"
    tower = None
    if want_symm:
        try:
            tower = DSSMTower(conf, max_nnz=max_nnz, device=env.dev, params=params, symmetric=True)
            ok = True
        except Exception as e:  # no symmetric-memory allocator on this box / torch build
            print(f"[bench] rank {env.rank}: symmetric memory unavailable ({type(e).__name__}: {e})", file=sys.stderr)
            ok = False
        if not env.all_true(ok):
            if env.rank == 0:
                print("[bench] falling back to --dp-comm nccl on ALL ranks", file=sys.stderr)
            dp_comm, tower = "nccl", None
    if tower is None:
        tower = DSSMTower(conf, max_nnz=max_nnz, device=env.dev, params=params)
    dp = DataParallelTower(tower, comm=dp_comm, sync_bn=sync_bn) if env.world > 1 else None
    return tower, dp


def timed_steps(env, step, steps, warmup, sampler=None):
    """W warm-up steps, then EXACTLY `steps` steps between barrier + synchronize, one CUDA event after every step on the
    launch stream: total = first -> last event (max over ranks), plus the per-step distribution of this rank."""
    torch = env.torch
    for i in range(warmup):
        step(i)
    env.barrier()
    if sampler is not None:
        sampler.start()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    env.barrier()
    evs[0].record(env.stream)
    for i in range(steps):
        step(i)
        evs[i + 1].record(env.stream)
    env.barrier()
    ms_total = env.max_over_ranks(evs[0].elapsed_time(evs[-1]))
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return ms_total, per


def run_train_workload(env, args, name, steps, warmup, primary):
    """Device-resident `value` leg and host-fed `e2e` leg of one workload; roofline block when primary."""
    import torch

    from dssm_b200.loader import HostBatchLoader
    from dssm_b200.synthetic import init_params, make_batch

    rank, world = env.rank, env.world
    conf = workload_conf(name, args.gemm_mode)
    NB = 4 if conf.rows > 20000 else 8  # distinct batches cycled through (each rank its own seeds)
    batches = [make_batch(conf, seed=1000 * rank + s) for s in range(NB)]
    mean_nnz = float(np.mean([b.nnz for b in batches]))
    params = init_params(conf, 0)
    tower, dp = build_tower(env, conf, batches, params, args.dp_comm)
    dev_batches = [tower.to_device(b) for b in batches]

    # ---- leg 1: device-resident inputs ("value") ------------------------------------------------------
    if world == 1:
        tower.capture_graph()

        def step(i):
            tower.stage(dev_batches[i % NB])  # D2D copy of the batch into the staging CSR
            tower.train_step_staged()  # CUDA-graph replay of fwd+bwd+Adam
    else:
        dp.capture_graph()

        def step(i):
            dp.train_step(dev_batches[i % NB])  # stage (D2D) + one graph: fwd, bwd, CSC, gather, exchange, Adam

    sampler = ClockSampler(env.local_rank) if (primary and rank == 0) else None
    launches0 = tower.launch_count
    replays0 = dp.replays if dp else 0
    ms_total, per = timed_steps(env, step, steps, warmup, sampler)
    # kernels of libdssm_b200.so inside the timed region; steps replayed as a torch-captured whole-step graph (N>1) are
    # counted as replays x the library launches of one eager step
    launches = tower.launch_count - launches0 + ((dp.replays - replays0) * dp.launches_per_step if dp else 0)
    clocks = sampler.stop() if sampler is not None else None
    ms_per_step = ms_total / steps
    value = conf.query_BS * world / (ms_per_step / 1e3)
    out = {"dp_comm": (dp.comm if dp else None), "dp_mc": (bool(getattr(dp, "use_multicast", False)) if dp else None),
           "conf": conf, "tower": tower, "dp": dp, "batches": batches, "params": params, "mean_nnz": mean_nnz, "NB": NB,
           "value": value, "ms_per_step": ms_per_step, "launches": int(launches), "clocks": clocks,
           "step_ms": {"median": pct(per, 50), "p10": pct(per, 10), "p90": pct(per, 90), "n": len(per),
                       "note": "per-step CUDA-event intervals of rank 0 inside the timed region"}}

    # ---- leg 2: end to end from HOST data through the public API ("e2e") ------------------------------
    # The epoch matrices (what the reference's training loop holds: query_train_dat / doc_train_dat / doc_neg_train_dat,
    # new_dssm.py:37-39) live in host memory; HostBatchLoader's worker thread assembles every step's stacked CSR into a
    # pinned ring INSIDE the timed region, the tower uploads it on its copy stream under the previous step's kernels
    # (dssm_tower_train_step_host_async / DataParallelTower.train_step_host_async) and every step's loss is read back by
    # the host, one step late so that the GPU queue never drains.
    q, p, n = epoch_matrices(batches, conf)
    loader = HostBatchLoader(q, p, n, conf.query_BS, conf.NEG, max_nnz=tower.max_nnz)
    step_async = tower.train_step_host_async if world == 1 else dp.train_step_host_async

    def run_epoch(n_steps):
        prev, last = None, None
        for pinned in loader.iterate([i % NB for i in range(n_steps)]):
            k = step_async(pinned)
            if prev is not None:
                last = tower.feed_loss(prev)
            prev = k
        if prev is not None:
            last = tower.feed_loss(prev)  # the last step's loss has landed on the host before the clock stops
        return last

    run_epoch(max(warmup, 2))
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(env.stream)
    last_loss = run_epoch(steps)
    e1.record(env.stream)
    env.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    e2e_ms = env.max_over_ranks(max(e0.elapsed_time(e1), wall_ms)) / steps
    out["e2e"] = {"value": conf.query_BS * world / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(4 * (conf.rows + 1) + 8 * mean_nnz),
                  "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms, "last_loss": last_loss,
                  "feed": "HostBatchLoader over host epoch matrices (batch assembly inside the clock) -> pinned ring -> double-buffered upload"}
    return out


def roofline_block(env, args, res, hbm_peak, peak_src, peaks):
    """Per-phase device time inside real (un-graphed) steps, CUDA events on the launch stream (dssm_tower_profile_step)."""
    tower, conf, dev_batches_n = res["tower"], res["conf"], res["NB"]
    dev_batches = [tower.to_device(b) for b in res["batches"]]
    acc = {}
    nprof = 6
    for i in range(nprof + 2):
        tower.stage(dev_batches[i % dev_batches_n])
        ph = tower.profile_step()
        if i >= 2:
            for k, v in ph.items():
                acc[k] = acc.get(k, 0.0) + v / nprof
    ab = algorithmic_bytes(conf, res["mean_nnz"], tower.P)
    kern = {k: {"ms": acc[k], "algorithmic_bytes": ab[k], "GBps": ab[k] / (acc[k] * 1e-3) / 1e9,
                "frac": ab[k] / (acc[k] * 1e-3) / 1e9 / hbm_peak} for k in ("spmm_fwd", "dw_gather", "adam")}
    dom = "spmm_fwd"  # the kernel north_star puts the >=70 %-of-HBM bar on
    traffic, l2_bytes = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath)).get(args.workload, {})
            traffic, l2_bytes = tj.get(dom), tj.get(dom + "_l2_bytes")
        except Exception:
            traffic = None
    dom_s = kern[dom]["ms"] * 1e-3
    step_bytes = float(sum(ab.values()))
    step_s = res["ms_per_step"] * 1e-3
    fl = dense_flops(conf)
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1426.5))
    dense_ms = acc["dense_fwd"] + acc["dense_bwd"]
    return {"kernel": "spmm_fwd_v4_kernel (FC1 CSR gather-accumulate)", "bound": "hbm", "achieved": kern[dom]["GBps"],
            "peak": hbm_peak, "unit": "GB/s", "frac": kern[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ab[dom], "avg_launch_ms": kern[dom]["ms"],
            # the contract `frac` charges one W1 row per non-zero; Zipf-hot rows and the 59 MB W1 live in L1/L2, so it
            # can exceed 1.  What the kernel really pulls from DRAM / L2 (ncu capture, per launch) over the live launch time:
            "dram_frac": (traffic / dom_s / 1e9 / hbm_peak) if traffic else None,
            "l2_frac_of_algorithmic": (l2_bytes / ab[dom]) if l2_bytes else None,
            "reading": "frac = SURVEY 8(d) algorithmic bytes / live launch time / HBM peak (a gather figure: > 1 means cache hits); "
                       "dram_frac = ncu dram bytes per launch / live launch time / HBM peak; the whole-step figure is in `step`",
            "step": {"algorithmic_bytes": step_bytes, "ms": res["ms_per_step"], "GBps": step_bytes / step_s / 1e9,
                     "frac": step_bytes / step_s / 1e9 / hbm_peak,
                     "roofline_ms": step_bytes / (hbm_peak * 1e9) * 1e3,
                     "terms": {k: float(v) for k, v in ab.items()}},
            "dense": {"flops_useful": fl, "ms_in_profiled_step": dense_ms, "TFLOPs_useful": fl / (dense_ms * 1e-3) / 1e12,
                      "tensor_issue_factor": 3 if conf.gemm_mode == "tc_3xtf32" else 1,
                      "frac_of_bf16_sustained_peak": fl / (dense_ms * 1e-3) / 1e12 / tf_peak,
                      "note": "dense_fwd + dense_bwd phases include the BN kernels; per-kernel tensor-pipe % is in profiles/ (ncu)"},
            "other_kernels": {k: v for k, v in kern.items() if k != dom},
            "phase_ms": acc, "phase_share_of_step": {k: v / sum(acc.values()) for k, v in acc.items()}}


def dp_parity(env, args, res, sync_bn=False, multicast=None):
    """Correctness evidence at N>1 where the driver can see it: restore known parameters, run ONE data-parallel step of the
    benchmarked workload and compare, on rank 0, against the oracle (DPOracle semantics: per-replica BN moments, mean
    gradient, one TF-Adam step; with sync_bn: the single-process reference on the re-stacked global batch); all ranks
    check that every replica's parameters are BIT-identical after the exchange."""
    import torch
    import torch.distributed as dist

    from dssm_b200.parallel import DataParallelTower
    from dssm_b200.synthetic import make_batch
    from oracle import DSSMOracle, OracleConfig

    tower, conf, params = res["tower"], res["conf"], res["params"]
    rank, world = env.rank, env.world
    t_start = time.perf_counter()
    dp = DataParallelTower(tower, comm=args.dp_comm if res["dp"].comm == "nvlink" else "nccl", sync_bn=sync_bn, multicast=multicast)
    if not sync_bn and tower.conf.use_bn:
        from dssm_b200._lib import check, lib

        check(lib.dssm_tower_set_syncbn(tower._h, 1, 0, None))
    # known state: initial parameters, zero optimizer state and shadows
    tower.load_params(params)
    tower.m.zero_(); tower.v.zero_(); tower.comm.zero_()
    tower.beta_pow.copy_(torch.tensor([conf.beta1, conf.beta2], dtype=torch.float32))
    b = make_batch(conf, seed=7000 + rank)
    assert b.nnz <= tower.max_nnz, (b.nnz, tower.max_nnz)
    loss_local = dp.train_step(tower.to_device(b)).item()
    torch.cuda.synchronize()
    # replicas bit-identical?
    ref = tower.params.clone()
    dist.broadcast(ref, src=0)
    identical = env.all_true(bool(torch.equal(ref, tower.params)))
    out = {"sync_bn": sync_bn, "comm": dp.comm, "nvls_multicast": bool(getattr(dp, "use_multicast", False)),
           "replicas_bit_identical": identical}
    losses = [torch.zeros(1, dtype=torch.float64, device=env.dev) for _ in range(world)]
    dist.all_gather(losses, torch.tensor([loss_local], dtype=torch.float64, device=env.dev))
    # relu active sets of every replica (sign of h*scale+shift, exact in float64): relu has a kink at 0, and a unit within
    # fp32 rounding of it gets derivative 1 in one implementation and 0 in another -- every gradient fed by that unit then
    # moves by its full upstream value, for ANY two fp32 implementations (about one unit per C2 step).  The oracle
    # therefore differentiates with the device's active set, after checking that the two sets differ only at units within
    # KINK_TOL of zero (tests/helpers.py:align_relu_masks does the same for the single-GPU parity tests).
    B, n_layers = conf.query_BS, len(conf.layers)
    masks = []
    for l in range(1, n_layers + 1):
        h = tower.tensor(f"h{l}").double()
        if conf.use_bn:
            sc, sh = tower.tensor(f"bn{l}_scale").double(), tower.tensor(f"bn{l}_shift").double()
            y = torch.cat([h[:B] * sc[0] + sh[0], h[B:] * sc[1] + sh[1]])
        else:
            y = h
        mine = (y > 0).to(torch.uint8).contiguous()
        allm = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        masks.append([m.cpu().numpy().astype(bool) for m in allm] if rank == 0 else None)
    KINK_TOL = 1e-4

    def align(cache, per_layer_masks):
        flips = 0
        if conf.act != "relu":
            return 0
        for l, m in enumerate(per_layer_masks, start=1):
            a = cache[f"a{l}"]
            diff = m != (a > 0)
            if diff.any():
                if float(np.abs(a[diff]).max()) >= KINK_TOL:
                    return -1  # active sets differ away from the kink: a real error
                flips += int(diff.sum())
                cache[f"a{l}"] = np.where(m, np.maximum(a, np.finfo(a.dtype).tiny), 0).astype(a.dtype)
        return flips

    if rank == 0:
        from threadpoolctl import threadpool_limits

        threads = host_threads()
        mats = [make_batch(conf, seed=7000 + r).to_scipy() for r in range(world)]

        def ocfg(B):
            return OracleConfig(TRIGRAM_D=conf.TRIGRAM_D, layers=tuple(conf.layers), NEG=conf.NEG, query_BS=B,
                                learning_rate=conf.learning_rate, use_bn=conf.use_bn, act=conf.act, loss_eps=conf.loss_eps,
                                loss_div_bs=conf.loss_div_bs)

        with threadpool_limits(limits=threads):
            if sync_bn:
                from oracle.syncbn import restack_global

                X = restack_global(mats, conf.query_BS, conf.NEG)
                o64 = threaded_port(ocfg(conf.query_BS * world), params, threads, np.float64)
                o32 = threaded_port(ocfg(conf.query_BS * world), params, threads)
                c64, c32 = o64.forward(X, on_train=True), o32.forward(X, on_train=True)
                gm = []
                for l in range(n_layers):  # global row order: every replica's queries, then positives, then negatives
                    parts = masks[l]
                    gm.append(np.concatenate([m[:B] for m in parts] + [m[B:2 * B] for m in parts] + [m[2 * B:] for m in parts]))
                kink_flips = align(c64, gm)
                align(c32, gm)
                g64, g32 = o64.backward(c64), o32.backward(c32)
                ref_loss = float(c64["loss"])
                got_loss = float(np.mean([l.item() for l in losses]))  # mean of the local losses = global-batch loss
                o32.adam_update(g32)
                ref_params = o32.p
            else:
                o64 = threaded_port(ocfg(conf.query_BS), params, threads, np.float64)
                o32 = threaded_port(ocfg(conf.query_BS), params, threads)
                g64 = g32 = None
                ref_loss = None
                kink_flips = 0
                for r, X in enumerate(mats):
                    c64, c32 = o64.forward(X, on_train=True, update_ema=False), o32.forward(X, on_train=True, update_ema=False)
                    f = align(c64, [masks[l][r] for l in range(n_layers)])
                    align(c32, [masks[l][r] for l in range(n_layers)])
                    kink_flips = -1 if (f < 0 or kink_flips < 0) else kink_flips + f
                    a, b32 = o64.backward(c64), o32.backward(c32)
                    if r == 0:
                        ref_loss = float(c64["loss"])
                    g64 = a if g64 is None else {k: g64[k] + a[k] for k in a}
                    g32 = b32 if g32 is None else {k: g32[k] + b32[k] for k in b32}
                g64 = {k: v / world for k, v in g64.items()}
                g32 = {k: (v / np.float32(world)).astype(np.float32) for k, v in g32.items()}
                got_loss = losses[0].item()
                o32.adam_update(g32)
                ref_params = o32.p
        loss_rel = abs(got_loss - ref_loss) / abs(ref_loss)
        got_g = tower.export_grads()
        worst, bad = 0.0, []
        for k, r64 in g64.items():
            if k == "W1" and dp.comm == "nvlink":
                continue  # the dense dW1 stays LOCAL under the fused exchange (only the owner ever sees the mean)
            if conf.use_bn and k[0] == "b" and k[1:].isdigit():
                continue  # analytically zero under BN: rounding noise on both sides
            scale = max(float(np.abs(r64).max()), 1e-30)
            allow = max(5e-5 * scale, 4 * float(np.abs(g32[k].astype(np.float64) - r64).max()))
            err = float(np.abs(got_g[k].astype(np.float64) - r64).max())
            worst = max(worst, err / scale)
            if err > allow:
                bad.append(k)
        got_p = tower.export_params()
        num = den = 0.0
        for k in ("W1", "W2", "W3"):
            if k in got_p:
                num += float(np.linalg.norm((got_p[k].astype(np.float64) - ref_params[k]).ravel()) ** 2)
                den += float(np.linalg.norm((ref_params[k].astype(np.float64) - params[k]).ravel()) ** 2)
        upd_rel = (num / max(den, 1e-300)) ** 0.5
        out.update({"loss_rel_err": loss_rel, "grad_max_rel_err": worst, "grads_outside_bounds": bad,
                    "relu_units_at_the_kink": kink_flips,
                    "param_update_rel_l2": upd_rel,
                    "bounds": "loss 1e-5; gradients max(5e-5 x scale, 4 x the fp32 port's own error vs float64); update rel-L2 0.1 "
                              "(first Adam step is lr*sign(g): near-zero gradients flip)",
                    "oracle_s": time.perf_counter() - t_start})
        out["ok"] = bool(identical and loss_rel <= 1e-5 and not bad and upd_rel <= 0.1 and kink_flips >= 0)
    dist.barrier()
    return out


def retrieval_block(env, args, peaks):
    """Corpus cosine top-k (BASELINE.json C5: 10 M x 128 docs over 8 GPUs, 4096 queries, top-100): every rank scores its
    1.25 M-doc shard (tcgen05 filter + exact rescoring), the (score, id) lists are all-gathered and merged on every rank.
    docs/s = corpus docs scored for the whole query batch per second (all ranks).  N>1: ids checked against the exact
    SIMT path on a query sub-sample over the same sharded corpus."""
    import torch
    import torch.distributed as dist

    from dssm_b200 import retrieval as rt

    dev, world, rank = env.dev, env.world, env.rank
    nqr, ndr, kr = 4096, 1_250_000, 100
    g = torch.Generator(device=dev).manual_seed(0)
    Qr = torch.relu(torch.randn((nqr, 128), generator=g, device=dev))  # same queries on every rank
    gd = torch.Generator(device=dev).manual_seed(100 + rank)
    Dr = torch.relu(torch.randn((ndr, 128), generator=gd, device=dev))  # this rank's shard
    total = ndr * world

    def once(Q, docs, method):
        if world == 1:
            return rt.corpus_topk(Q, docs, kr, method=method)
        return rt.sharded_corpus_topk(Q, docs, kr, id_offset=rank * ndr, total_docs=total, method=method)

    def timed(docs, method, reps=3):
        once(Qr, docs, method)  # warm-up (attribute calls, allocator, NCCL channels)
        env.barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(env.stream)
        for _ in range(reps):
            out = once(Qr, docs, method)
        r1.record(env.stream)
        env.barrier()
        return env.max_over_ranks(r0.elapsed_time(r1)) / reps, out, bool(rt.LAST_CALL["fallback"])

    # (a) bf16-stored corpus index (built once per corpus, 256 B per doc streamed per query batch) -- the headline variant;
    # (b) filter straight from the fp32 rows (512 B per doc, no index).  Both re-score survivors exactly from the fp32 rows.
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(env.stream)
    index = rt.CorpusIndex(Dr)
    b1.record(env.stream)
    torch.cuda.synchronize()
    build_ms = b0.elapsed_time(b1)
    ms_bf16, (rs_, ri_), fb_bf16 = timed(index, "bf16")
    ms_fp32, (fs_, fi_), fb_fp32 = timed(Dr, "tc")
    same_variants = env.all_true(bool(torch.equal(ri_, fi_) and torch.equal(rs_, fs_)))
    check = None
    if world > 1:
        sub = 64
        es, ei = once(Qr[:sub].contiguous(), Dr, "exact")
        same = bool(torch.equal(ei, ri_[:sub]) and torch.equal(es, rs_[:sub]))
        check = {"queries_checked": sub, "ids_and_scores_equal_exact_path": env.all_true(same)}
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1426.5))
    ceiling = world * tf_peak * 1e12 / (2.0 * nqr * 128)  # docs/s if the dense bf16 contraction ran at the measured cuBLAS rate

    def rec(ms, fb):
        v = total / (ms / 1e3)
        return {"value": v, "ms": ms, "tensor_tflops": 2.0 * nqr * total * 128 / (ms / 1e3) / 1e12, "frac_of_bf16_tensor_ceiling": v / ceiling,
                "fallback_to_exact": fb}

    head = rec(ms_bf16, fb_bf16)
    return {"metric": "corpus cos top-k docs/s", "value": head["value"], "unit": "docs/s", "ms": ms_bf16, "n_gpus": world,
            "config": {"queries": nqr, "docs_total": total, "docs_per_gpu": ndr, "dim": 128, "k": kr, "storage": "bf16 index + fp32 rows",
                       "method": "tcgen05 bf16 filter over a per-corpus index (normalised rows as swizzled tiles) + exact fp32 rescoring "
                                 "(ids and scores bit-exact vs oracle); N>1: NCCL all-gather of the per-shard lists + merge kernel on every rank"},
            "fallback_to_exact": fb_bf16, "tensor_tflops": head["tensor_tflops"], "frac_of_tensor_ceiling": head["frac_of_bf16_tensor_ceiling"],
            "index_build_ms": build_ms, "index_bytes_per_gpu": index.nbytes,
            "variants": {"bf16_index": head, "fp32_rows_tf32_filter": rec(ms_fp32, fb_fp32), "identical_ids_and_scores": same_variants},
            "sharded_check": check,
            "includes": "query norms + query image, exact seed pass, filter passes, rescoring, overflow-flag readback" +
                        (", all-gather, merge" if world > 1 else "") + "; the one-time index build is reported separately"}


def run_ours(args):
    # NCCL prints its version banner on stdout: keep fd 1 clean for the single JSON line
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    env = Env(args)
    import torch
    import torch.distributed as dist

    rank, world = env.rank, env.world
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"

    res = run_train_workload(env, args, args.workload, args.steps, args.warmup, primary=True)
    conf, tower, dp = res["conf"], res["tower"], res["dp"]
    roof = roofline_block(env, args, res, hbm_peak, peak_src, peaks) if rank == 0 and world == 1 else None
    if world > 1:
        dist.barrier()

    # ---- correctness evidence for the multi-GPU paths (N>1) -----------------------------------------------------
    parity = None
    if world > 1 and not args.no_dp_parity:
        dp.graph = None
        parity = {"default": dp_parity(env, args, res)}
        if dp.comm == "nvlink":
            other = not bool(getattr(dp, "use_multicast", False))
            parity["nvls_multicast" if other else "peer_loads_stores"] = dp_parity(env, args, res, multicast=other)
        if conf.use_bn:
            parity["sync_bn"] = dp_parity(env, args, res, sync_bn=True)
        if rank == 0:
            parity["ok"] = all(v.get("ok", False) for v in parity.values() if isinstance(v, dict))

    # ---- single-pass tf32 dense mode (stated tolerance: include/dssm_b200.h) --------------------------------------
    tf32 = None
    if world == 1 and not args.no_extras and args.gemm_mode == "tc_3xtf32":
        a2 = argparse.Namespace(**vars(args))
        a2.gemm_mode = "tc_tf32"
        r2 = run_train_workload(env, a2, args.workload, min(args.steps, 20), 3, primary=False)
        tf32 = {"gemm_mode": "tc_tf32", "value": r2["value"], "ms_per_step": r2["ms_per_step"],
                "tolerance": "loss / embeddings / cosines within 5e-3 of scale vs the float64 oracle (tests/test_gpu_tower.py::test_tf32_mode_tolerance)"}
        del r2

    # ---- BASELINE configs[2]: C3 = 8192 groups per GPU, at every N --------------------------------------------------
    c3 = None
    if not args.no_extras and args.workload.upper() != "C3":
        res["tower"] = res["dp"] = tower = dp = None
        torch.cuda.empty_cache()
        r3 = run_train_workload(env, args, "C3", min(args.steps, 10), 3, primary=False)
        c3 = {"workload": describe(r3["conf"], "C3", world)["workload"], "value": r3["value"], "unit": UNIT, "n_gpus": world,
              "ms_per_step": r3["ms_per_step"], "steps": min(args.steps, 10), "step_ms": r3["step_ms"], "e2e": r3["e2e"],
              "dp_comm": (r3["dp"].comm if r3["dp"] else None),
              "dp_nvls_multicast": (getattr(r3["dp"], "use_multicast", None) if r3["dp"] else None)}
        if r3["dp"] is not None:
            r3["dp"].graph = None
        del r3
        torch.cuda.empty_cache()

    # ---- secondary metric: corpus cosine top-k (BASELINE.json C5) -----------------------------------------------------
    retrieval = None
    if not args.no_retrieval:
        retrieval = retrieval_block(env, args, peaks)

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores -----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        nsteps = 6 if conf.query_BS <= 1024 else 2
        from threadpoolctl import threadpool_limits

        with threadpool_limits(limits=threads):
            times = cpu_port_step_time(conf, res["batches"][:2], res["params"], nsteps, threads)
        cpu = {"value": conf.query_BS / float(np.mean(times)), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{nsteps} full {args.workload} train steps (query_BS={conf.query_BS}) of the NumPy/SciPy oracle "
                         f"(sparse products, Adam and dense layers on all {threads} host threads)",
               "ms_per_step": 1e3 * float(np.mean(times))}

    if rank == 0:
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": describe(conf, args.workload, world),
                "run": {"mean_nnz_per_step_per_gpu": res["mean_nnz"], "gemm_mode": conf.gemm_mode, "distinct_batches": res["NB"],
                        "cuda_graph": True, "dp_comm": res["dp_comm"], "dp_nvls_multicast": res["dp_mc"],
                        "bn_moments": "per-replica (the SyncBN option is exercised in dp_parity)" if world > 1 else "batch",
                        "deterministic": "bit-reproducible step (row-ordered CSC columns, fixed-order merges)" +
                                         ("; NVLS multimem.ld_reduce order is the switch's" if res["dp_mc"] else "")},
                "step_ms": res["step_ms"], "e2e": res["e2e"], "gpu_launches": res["launches"], "clocks": res["clocks"],
                "roofline": roof, "cpu_baseline": cpu, "dp_parity": parity, "tf32_single_pass": tf32, "c3": c3, "retrieval": retrieval}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        # Tear-down: a process group that has collectives captured in a live CUDA graph can block forever in
        # destroy_process_group (seen on NCCL 2.28.9: the JSON line was out, the ranks never exited).  Drop the
        # graph first, meet at a barrier, and leave without the NCCL destructor if it does not return promptly.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        threading.Timer(10.0, lambda: os._exit(0)).start()
        try:
            dist.destroy_process_group()
        finally:
            os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", help="C1 | C2 | C3 | C4 | C4_NOBN (per-GPU batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-retrieval", action="store_true", help="skip the secondary corpus top-k measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the C3 block and the single-pass tf32 leg")
    ap.add_argument("--no-dp-parity", action="store_true", help="N>1: skip the oracle check of one data-parallel step")
    ap.add_argument("--dp-comm", default="nvlink", choices=["nccl", "nvlink"],
                    help="N>1: dW1 exchange by NCCL all-reduce (chunked, overlapped) or by the fused pull/Adam/push kernel over "
                         "NVLink peer memory (csrc/nvlink.cu)")
    ap.add_argument("--gemm-mode", default="tc_3xtf32", choices=["fp32", "tc_3xtf32", "tc_tf32"],
                    help="dense-layer arithmetic: FFMA fp32 or tcgen05 3xTF32 (both hold the 1e-5 parity bar), or single-pass tf32")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20  # bounded sample: the CPU port needs ~0.8 s per C2 step
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
