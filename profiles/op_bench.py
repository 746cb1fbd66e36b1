"""Warm-cache duration of the individual kernels of the dense stack at a config's shapes: G launches of one op are
captured into a CUDA graph and replayed, so the figure is kernel time + the ~1 us node-to-node gap, with the operands
L2-resident as they are inside the real step.   python profiles/op_bench.py [C2|C3|C4] [3|1]"""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from dssm_b200 import baseline_config
from dssm_b200._lib import lib, check
from dssm_b200.synthetic import make_batch
from dssm_b200 import ops

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
PASSES = int(sys.argv[2]) if len(sys.argv) > 2 else 3  # 3 = 3xTF32 (parity mode), 1 = single-pass tf32
conf = baseline_config(name)
B, R = conf.query_BS, (2 + conf.NEG) * conf.query_BS
dev = torch.device("cuda")
G, REP = 20, 10
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
f32 = lambda *s: torch.randn(*s, device=dev, dtype=torch.float32)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def bench(label, fn, bytes_moved=None):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(G):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REP):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (G * REP)
    extra = "" if bytes_moved is None else f"   {bytes_moved / us / 1e3:7.0f} GB/s algorithmic"
    print(f"{name} {label:42s} {us:8.2f} us{extra}", flush=True)


for fn_name, args in (("dssm_fc_tc_image_bytes", None), ("dssm_fc_tc_build_image", None), ("dssm_fc_fwd_tc_img", None), ("dssm_fc_bwd_dx_tc_img", None)):
    getattr(lib, fn_name).restype = C.c_size_t if fn_name.endswith("bytes") else C.c_int
i32, vp = C.c_int32, C.c_void_p
lib.dssm_fc_tc_image_bytes.argtypes = [i32, i32, i32, i32]
lib.dssm_fc_tc_build_image.argtypes = [vp, i32, i32, i32, i32, vp, vp]
lib.dssm_fc_fwd_tc_img.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, vp, i32, vp, i32, vp, vp]
lib.dssm_fc_bwd_dx_tc_img.argtypes = [vp, i32, i32, vp, i32, vp, i32, vp]

dims = [conf.TRIGRAM_D] + list(conf.layers)
for l in range(2, len(dims)):
    K, N = dims[l - 1], dims[l]
    H, W, b = f32(R, K), f32(K, N) * 0.05, f32(N)
    sc, sh = f32(2, K).abs() + 0.5, f32(2, K)
    out, dH, dA = f32(R, N), f32(R, N), f32(R, K)
    imgf = torch.zeros(lib.dssm_fc_tc_image_bytes(K, N, 0, R), dtype=torch.uint8, device=dev)
    imgx = torch.zeros(lib.dssm_fc_tc_image_bytes(K, N, 1, R), dtype=torch.uint8, device=dev)
    check(lib.dssm_fc_tc_build_image(p(W), K, N, 0, R, p(imgf), st()))
    check(lib.dssm_fc_tc_build_image(p(W), K, N, 1, R, p(imgx), st()))
    # fused BN column moments in the epilogue (FusedBnStats, bn_common.cuh): host struct built with ctypes
    class BnFin(C.Structure):
        _fields_ = [(n_, C.c_void_p) for n_ in ("gamma", "beta", "ema_mean", "ema_var", "mean", "var", "rstd", "scale", "shift")] + \
                   [("eps", C.c_float), ("decay", C.c_float), ("update_ema", C.c_int), ("nq_chunks", C.c_int)]

    class Fbn(C.Structure):
        _fields_ = [("part", C.c_void_p), ("tickets", C.c_void_p), ("fin", BnFin), ("on", C.c_int)]

    if B % 128 == 0 and R <= 128 * 64:  # the fused-moments finalize stages at most 64 M tiles
        z2 = lambda: f32(2, N)
        keep = [z2() for _ in range(9)]
        partb = torch.zeros(3 * ((R + 127) // 128) * N, device=dev)
        tick = torch.zeros(64, dtype=torch.int32, device=dev)
        for bits, lab in ((1, "fused BN moments"), (1 | 2, "  .. no finalize"), (1 | 4, "  .. no tile partial / ticket / finalize"), (1 | 8 | 4, "  .. nothing (flag only)")):
            fb = Fbn(p(partb).value, p(tick).value, BnFin(*[p(k_).value for k_ in keep], 1e-3, 0.5, 1, B // 128), bits)
            bench(f"fc_fwd  tc img  + {lab}", lambda: check(lib.dssm_fc_fwd_tc_img(p(H), R, K, B, p(sc), p(sh), 1, p(imgf), p(b), N, p(out), PASSES, C.byref(fb), st())),
                  4 * R * (K + N))
    bench(f"fc_fwd  tc img  [{R}x{K}]x[{K}x{N}]", lambda: check(lib.dssm_fc_fwd_tc_img(p(H), R, K, B, p(sc), p(sh), 1, p(imgf), p(b), N, p(out), PASSES, None, st())),
          4 * R * (K + N))
    bench(f"fc_dx   tc img  [{R}x{N}]x[{N}x{K}]", lambda: check(lib.dssm_fc_bwd_dx_tc_img(p(dH), R, N, p(imgx), K, p(dA), PASSES, st())), 4 * R * (K + N))
    ws = torch.zeros(lib.dssm_fc_bwd_dw_workspace_bytes(R, K, N), dtype=torch.uint8, device=dev)
    dW, db = f32(K, N), f32(N)
    bench(f"fc_dw   tc      [{K}x{R}]x[{R}x{N}] (+reduce)",
          lambda: check(lib.dssm_fc_bwd_dw(p(H), R, K, B, p(sc), p(sh), 1, p(dH), N, p(dW), None, 1 if PASSES == 3 else 2, p(ws), ws.numel(), st())), 4 * R * (K + N))
    bench(f"image build x1  [{K}x{N}]", lambda: check(lib.dssm_fc_tc_build_image(p(W), K, N, 0, R, p(imgf), st())))
for L in sorted(set(conf.layers)):
    X, dAl = f32(R, L), f32(R, L)
    state = ops.BNState(L, dev)
    ws = torch.zeros(lib.dssm_bn_workspace_bytes(R, L), dtype=torch.uint8, device=dev)
    fwd = lambda: check(lib.dssm_bn_forward(p(X), R, L, B, 1, 1, p(state.gamma), p(state.beta), p(state.ema_mean), p(state.ema_var), 1e-3, 0.5,
                                            p(state.mean), p(state.var), p(state.rstd), p(state.scale), p(state.shift), p(ws), ws.numel(), st()))
    bench(f"bn_forward (stats+finalize) [{R}x{L}]", fwd, 4 * R * L)
    dg, dbt, dbb = f32(2, L), f32(2, L), f32(L)
    bench(f"bn_act_backward (reduce+apply) [{R}x{L}]",
          lambda: check(lib.dssm_bn_act_backward(p(dAl), p(X), R, L, B, 1, p(state.gamma), p(state.mean), p(state.rstd), p(state.scale), p(state.shift),
                                                 p(dg), p(dbt), p(dbb), p(ws), ws.numel(), st())), 4 * R * L * 4)
L = conf.layers[-1]
Hl, Y, dY = f32(R, L), f32(R, L), f32(R, L)
state = ops.BNState(L, dev); state.scale.fill_(1.0)
K1 = 1 + conf.NEG
qn, dn, cr, cs, pr, lt, ls = f32(B), f32(K1 * B), f32(K1 * B), f32(B, K1), f32(B, K1), f32(B), f32(4)
bench(f"cos_softmax_loss_fused [{R}x{L}]",
      lambda: check(lib.dssm_cos_softmax_loss_fused(p(Hl), p(state.scale), p(state.shift), 1, p(Y), B, conf.NEG, L, 20.0, 0.0, 1, p(qn), p(dn), p(cr),
                                                    p(cs), p(pr), p(lt), p(ls), p(dY), st())), 4 * R * L * 3)
bt = make_batch(conf, seed=1)
x = ops.DeviceCSR.from_host(bt, dev)
if x is not None:
    W1, b1, h1 = f32(conf.TRIGRAM_D, dims[1]) * 0.01, f32(dims[1]), f32(R, dims[1])
    bench(f"spmm_fwd nnz={bt.nnz}", lambda: check(lib.dssm_spmm_fwd(p(x.indptr), p(x.indices), p(x.values), R, conf.TRIGRAM_D, p(W1), p(b1), dims[1], p(h1), st())),
          bt.nnz * (8 + 4 * dims[1]) + 4 * R * dims[1])
