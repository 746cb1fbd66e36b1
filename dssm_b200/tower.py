"""DSSMTower: the reference graph (semantic_matching/dssm/new_dssm.py:104-217) as a drop-in tower.

The reference drives its graph with ``sess.run(fetch, feed_dict=pull_batch(...))`` (new_dssm.py:261-286) and
looks tensors up by graph name (load_model_and_save_vector.py:30-46).  ``DSSMTower.run`` keeps that call
shape; ``train_step`` / ``forward`` are the direct forms.  All arithmetic happens in libdssm_b200.so -- one C
call per step; torch owns the device memory and the stream, nothing else.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Union

import numpy as np
import torch

from . import _lib
from ._lib import ACT, GEMM, DssmError, check, last_error, lib, ptr, stream_ptr
from .batch import DOC_NEG_BATCH, DOC_POS_BATCH, ON_TRAIN, QUERY_BATCH, StackedBatch, stack_feed
from .config import Config
from .ops import DeviceCSR
from .synthetic import init_params

# reference graph names -> tower tensors (new_dssm.py:156-158,185-199,206-209; with ":0" or scope prefix accepted)
_ALIASES = {
    "embedding_query_y": "embedding_query_y",
    "embedding_doc_positive_y": "embedding_doc_positive_y",
    "embedding_doc_negative_y": "embedding_doc_negative_y",
    "query_norm_single": "query_norm_single",
    "doc_norm": "doc_norm",
    "cos_sim_raw": "cos_sim_raw",
    "cos_sim": "cos_sim",
    "prob": "prob",
    "hit_prob": "hit_prob",
    "loss": "loss",
    "accuracy": "accuracy",
}


def _make_c_config(conf: Config) -> _lib.dssm_config:
    c = _lib.dssm_config()
    c.TRIGRAM_D, c.n_layers = conf.TRIGRAM_D, len(conf.layers)
    for i, n in enumerate(conf.layers):
        c.layers[i] = n
    c.NEG, c.query_BS = conf.NEG, conf.query_BS
    c.use_bn, c.act = int(conf.use_bn), ACT[conf.act]
    c.loss_div_bs, c.gemm_mode = int(conf.loss_div_bs), GEMM[conf.gemm_mode]
    c.bn_eps, c.ema_decay, c.gamma, c.loss_eps = conf.bn_eps, conf.ema_decay, conf.gamma, conf.loss_eps
    c.learning_rate, c.beta1, c.beta2, c.adam_eps = conf.learning_rate, conf.beta1, conf.beta2, conf.adam_eps
    return c


class DSSMTower:
    def __init__(self, conf: Config, max_nnz: int, device: Union[str, torch.device] = "cuda",
                 params: Optional[Dict[str, np.ndarray]] = None, seed: int = 0, symmetric: bool = False):
        if len(conf.layers) > _lib.MAX_LAYERS:
            raise ValueError("too many layers")
        self.conf = conf
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("DSSMTower needs a CUDA device (there is no CPU path)")
        self.max_nnz = int(max_nnz)
        self._h = C.c_void_p()
        cc = _make_c_config(conf)
        check(lib.dssm_tower_create(C.byref(cc), C.byref(self._h)))
        self.P = lib.dssm_tower_param_count(self._h)
        self.E = lib.dssm_tower_ema_count(self._h)
        with torch.cuda.device(self.device):
            z = lambda n: torch.zeros(max(int(n), 4), dtype=torch.float32, device=self.device)
            self.symmetric = bool(symmetric)
            if symmetric:
                # parameters and gradients in peer-mappable (symmetric) memory: every rank of a process group can address
                # them over NVLink once rendezvous'ed (DataParallelTower(comm="nvlink")); needs torch.distributed
                import torch.distributed._symmetric_memory as symm_mem

                zs = lambda n: symm_mem.empty(max(int(n), 4), dtype=torch.float32, device=self.device).zero_()
            else:
                zs = z
            self.params, self.m, self.v = zs(self.P), z(self.P), z(self.P)
            # grads and the EMA shadows share one allocation so that data-parallel training exchanges
            # [small gradients | EMA] with a single collective (dssm_b200/parallel.py)
            self.comm = zs(self.P + max(self.E, 4))
            self.grads, self.ema = self.comm[:self.P], self.comm[self.P:self.P + max(self.E, 4)]
            self.beta_pow = torch.tensor([conf.beta1, conf.beta2], dtype=torch.float32, device=self.device)
            ws_bytes = lib.dssm_tower_workspace_bytes(self._h, self.max_nnz)
            self.workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=self.device)
            check(lib.dssm_tower_bind(self._h, ptr(self.params), ptr(self.grads), ptr(self.m), ptr(self.v), ptr(self.ema),
                                      ptr(self.beta_pow), ptr(self.workspace), ws_bytes, self.max_nnz))
        self._layout = {k: self._tensor_table(k) for k in (0, 1, 2)}
        self._ws_f32 = self.workspace.view(torch.float32)
        self.load_params(params if params is not None else init_params(conf, seed))
        self._graph = False
        self._pinned = None
        self._host_loss = torch.zeros(1, dtype=torch.float32).pin_memory() if torch.cuda.is_available() else None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                lib.dssm_tower_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- layout ---------------------------------------------------------------------------------------
    def _tensor_table(self, kind: int):
        out = {}
        name = C.create_string_buffer(64)
        off, rows, cols = C.c_int64(), C.c_int64(), C.c_int64()
        for i in range(lib.dssm_tower_num_tensors(self._h, kind)):
            check(lib.dssm_tower_tensor_info(self._h, kind, i, name, 64, C.byref(off), C.byref(rows), C.byref(cols)))
            out[name.value.decode()] = (off.value, rows.value, cols.value)
        return out

    def _view(self, buf: torch.Tensor, entry):
        off, rows, cols = entry
        return buf[off:off + rows * cols].view(rows, cols)

    def param(self, name: str, which: str = "params") -> torch.Tensor:
        """View of a named tensor inside the flat params / grads / m / v buffer (names: W{l}, b{l},
        bn{l}_gamma, bn{l}_beta with [2,L] = (query, doc) instances)."""
        return self._view(getattr(self, which), self._layout[0][name])

    def ema_tensor(self, name: str) -> torch.Tensor:
        return self._view(self.ema, self._layout[1][name])

    def tensor(self, name: str) -> torch.Tensor:
        """A tensor of the last forward/backward by graph name (reference names accepted, e.g.
        'BN2/embedding_query_y:0')."""
        key = name.split("/")[-1].split(":")[0]
        B = self.conf.query_BS
        if key in self._layout[2]:
            return self._view(self._ws_f32, self._layout[2][key])
        if key == "embedding_query_y":
            return self.tensor("Y")[:B]
        if key == "embedding_doc_positive_y":
            return self.tensor("Y")[B:2 * B]
        if key == "embedding_doc_negative_y":
            return self.tensor("Y")[2 * B:]
        if key == "hit_prob":
            return self.tensor("prob")[:, 0:1]
        if key == "accuracy":  # new_dssm.py:220-221, on the device (csrc/metrics.cu)
            if getattr(self, "_acc_out", None) is None:
                self._acc_out = torch.zeros(1, dtype=torch.float32, device=self.device)
            prob = self.tensor("prob")
            check(lib.dssm_accuracy(ptr(prob), prob.shape[0], prob.shape[1], ptr(self._acc_out), stream_ptr()))
            return self._acc_out[0]
        if key in self._layout[0]:
            return self.param(key)
        if key in self._layout[1]:
            return self.ema_tensor(key)
        raise KeyError(name)

    # ---- parameters -----------------------------------------------------------------------------------
    def load_params(self, p: Dict[str, np.ndarray]) -> None:
        """Accepts the oracle/synthetic naming: W{l}, b{l}, bn{l}_{q|d}_{beta|gamma}."""
        n = len(self.conf.layers)
        for l in range(1, n + 1):
            self.param(f"W{l}").copy_(torch.from_numpy(np.ascontiguousarray(p[f"W{l}"], dtype=np.float32)))
            self.param(f"b{l}").copy_(torch.from_numpy(np.ascontiguousarray(p[f"b{l}"], dtype=np.float32)).view(1, -1))
            if self.conf.use_bn:
                for which in ("gamma", "beta"):
                    stacked = np.stack([p[f"bn{l}_q_{which}"], p[f"bn{l}_d_{which}"]]).astype(np.float32)
                    self.param(f"bn{l}_{which}").copy_(torch.from_numpy(stacked))

    def _export(self, which: str) -> Dict[str, np.ndarray]:
        out = {}
        n = len(self.conf.layers)
        for l in range(1, n + 1):
            out[f"W{l}"] = self.param(f"W{l}", which).cpu().numpy().copy()
            out[f"b{l}"] = self.param(f"b{l}", which).cpu().numpy().reshape(-1).copy()
            if self.conf.use_bn:
                for w in ("gamma", "beta"):
                    t = self.param(f"bn{l}_{w}", which).cpu().numpy()
                    out[f"bn{l}_q_{w}"], out[f"bn{l}_d_{w}"] = t[0].copy(), t[1].copy()
        return out

    def export_params(self):
        return self._export("params")

    def export_grads(self):
        return self._export("grads")

    def export_ema(self) -> Dict[str, np.ndarray]:
        out = {}
        if self.conf.use_bn:
            for l in range(1, len(self.conf.layers) + 1):
                for nm in ("mean", "var"):
                    t = self.ema_tensor(f"bn{l}_ema_{nm}").cpu().numpy()
                    out[f"bn{l}_q_ema_{nm}"], out[f"bn{l}_d_ema_{nm}"] = t[0].copy(), t[1].copy()
        return out

    def state_dict(self) -> Dict[str, np.ndarray]:
        """Everything tf.train.Saver() would hold (new_dssm.py:248): trainables, EMA shadows, Adam slots."""
        sd = {f"params/{k}": v for k, v in self._export("params").items()}
        sd.update({f"adam_m/{k}": v for k, v in self._export("m").items()})
        sd.update({f"adam_v/{k}": v for k, v in self._export("v").items()})
        sd.update({f"ema/{k}": v for k, v in self.export_ema().items()})
        sd["beta_pow"] = self.beta_pow.cpu().numpy().copy()
        return sd

    # ---- on-disk checkpoint (the reference: tf.train.Saver().save(sess, "model/model_1.ckpt"), new_dssm.py:248,331;
    #      vocabulary pickled beside it, utils/utils.py:241-261) ------------------------------------------------------
    CKPT_FORMAT = "dssm_b200.ckpt.v1"

    def save(self, path: str, vocabulary: Optional[Dict[str, int]] = None, state: Optional[Dict[str, np.ndarray]] = None) -> str:
        """Write everything tf.train.Saver() holds for this graph -- the 12 trainables (W{l}, b{l}, bn{l}_{q|d}_{beta|gamma}),
        the 8 EMA shadows, the Adam slots m / v of every trainable and the two beta powers -- plus the hyper-parameters and,
        optionally, the fitted vocabulary (token -> column, what the reference pickles as output/vectorizer_data) into ONE
        .npz file (TF's V2 checkpoint files cannot be written without TF; the variable SET is the reference's).  Returns the
        path written.  `state`: a state dict to write instead of this tower's (DataParallelTower.state_dict())."""
        import json

        sd = dict(state if state is not None else self.state_dict())
        c = self.conf
        meta = {"format": self.CKPT_FORMAT, "TRIGRAM_D": c.TRIGRAM_D, "layers": list(c.layers), "NEG": c.NEG, "query_BS": c.query_BS,
                "learning_rate": c.learning_rate, "use_bn": c.use_bn, "act": c.act, "loss_eps": c.loss_eps, "loss_div_bs": c.loss_div_bs,
                "bn_eps": c.bn_eps, "ema_decay": c.ema_decay, "gamma": c.gamma, "beta1": c.beta1, "beta2": c.beta2, "adam_eps": c.adam_eps}
        sd["__meta__"] = np.frombuffer(json.dumps(meta).encode("utf-8"), dtype=np.uint8)
        if vocabulary is not None:
            toks = sorted(vocabulary, key=lambda k: vocabulary[k])
            sd["__vocab_tokens__"] = np.frombuffer("\n".join(toks).encode("utf-8"), dtype=np.uint8)
            sd["__vocab_ids__"] = np.asarray([vocabulary[k] for k in toks], dtype=np.int64)
        if not path.endswith(".npz"):
            path = path + ".npz"
        with open(path, "wb") as f:
            np.savez(f, **sd)
        return path

    @staticmethod
    def read_checkpoint(path: str):
        """(state dict, meta dict, vocabulary or None) of a file written by save()."""
        import json

        z = np.load(path if path.endswith(".npz") else path + ".npz")
        meta = json.loads(bytes(z["__meta__"]).decode("utf-8"))
        if meta.get("format") != DSSMTower.CKPT_FORMAT:
            raise ValueError(f"{path}: not a {DSSMTower.CKPT_FORMAT} checkpoint")
        vocab = None
        if "__vocab_tokens__" in z.files:
            toks = bytes(z["__vocab_tokens__"]).decode("utf-8").split("\n")
            vocab = {t: int(i) for t, i in zip(toks, z["__vocab_ids__"])}
        sd = {k: z[k] for k in z.files if not k.startswith("__")}
        return sd, meta, vocab

    def restore(self, path: str) -> Optional[Dict[str, int]]:
        """saver.restore(sess, ckpt) (load_model_and_save_vector.py:10-11): loads parameters, EMA shadows and optimizer state;
        refuses a checkpoint of a different graph.  Returns the stored vocabulary (or None)."""
        sd, meta, vocab = self.read_checkpoint(path)
        c = self.conf
        if int(meta["TRIGRAM_D"]) != c.TRIGRAM_D or tuple(meta["layers"]) != tuple(c.layers) or bool(meta["use_bn"]) != c.use_bn:
            raise ValueError(f"{path}: checkpoint of TRIGRAM_D={meta['TRIGRAM_D']} layers={meta['layers']} use_bn={meta['use_bn']} "
                             f"does not fit this tower (TRIGRAM_D={c.TRIGRAM_D} layers={list(c.layers)} use_bn={c.use_bn})")
        self.load_state_dict(sd)
        return vocab

    @classmethod
    def from_checkpoint(cls, path: str, max_nnz: int, device="cuda", query_BS: Optional[int] = None, NEG: Optional[int] = None,
                        gemm_mode: str = "fp32", **kw):
        """Build a tower of the checkpoint's graph (the import_meta_graph + restore pair of load_model_and_save_vector.py:10-11).
        query_BS / NEG may differ from training time: they only shape the batch."""
        sd, meta, vocab = cls.read_checkpoint(path)
        conf = Config(TRIGRAM_D=int(meta["TRIGRAM_D"]), query_BS=int(query_BS or meta["query_BS"]), NEG=int(NEG or meta["NEG"]),
                      learning_rate=float(meta["learning_rate"]), layers=tuple(meta["layers"]), use_bn=bool(meta["use_bn"]),
                      act=meta["act"], loss_eps=float(meta["loss_eps"]), loss_div_bs=bool(meta["loss_div_bs"]), gemm_mode=gemm_mode)
        t = cls(conf, max_nnz=max_nnz, device=device, **kw)
        t.load_state_dict(sd)
        t.vocabulary = vocab
        return t

    def load_state_dict(self, sd: Dict[str, np.ndarray]) -> None:
        self.load_params({k[len("params/"):]: v for k, v in sd.items() if k.startswith("params/")})
        n = len(self.conf.layers)
        for slot, which in (("adam_m", "m"), ("adam_v", "v")):
            for l in range(1, n + 1):
                self.param(f"W{l}", which).copy_(torch.from_numpy(sd[f"{slot}/W{l}"]))
                self.param(f"b{l}", which).copy_(torch.from_numpy(sd[f"{slot}/b{l}"]).view(1, -1))
                if self.conf.use_bn:
                    for w in ("gamma", "beta"):
                        st = np.stack([sd[f"{slot}/bn{l}_q_{w}"], sd[f"{slot}/bn{l}_d_{w}"]]).astype(np.float32)
                        self.param(f"bn{l}_{w}", which).copy_(torch.from_numpy(st))
        if self.conf.use_bn:
            for l in range(1, n + 1):
                for nm in ("mean", "var"):
                    st = np.stack([sd[f"ema/bn{l}_q_ema_{nm}"], sd[f"ema/bn{l}_d_ema_{nm}"]]).astype(np.float32)
                    self.ema_tensor(f"bn{l}_ema_{nm}").copy_(torch.from_numpy(st))
        self.beta_pow.copy_(torch.from_numpy(np.asarray(sd["beta_pow"], dtype=np.float32)))

    # ---- execution ------------------------------------------------------------------------------------
    def _check_batch(self, x: DeviceCSR):
        if x.rows != self.conf.rows:
            raise ValueError(f"batch has {x.rows} rows, the graph needs exactly (2+NEG)*query_BS = {self.conf.rows}")
        if x.n_cols != self.conf.TRIGRAM_D:
            raise ValueError("batch TRIGRAM_D mismatch")
        if x.nnz > self.max_nnz:
            raise ValueError(f"batch nnz {x.nnz} exceeds max_nnz {self.max_nnz} the tower was sized for")

    def to_device(self, b: StackedBatch) -> DeviceCSR:
        return DeviceCSR.from_host(b, self.device)

    def forward(self, x: DeviceCSR, on_train: bool = False, update_ema: Optional[bool] = None) -> torch.Tensor:
        """sess.run(loss, feed) -- new_dssm.py:276-278.  Returns the device loss scalar; other tensors via tensor()."""
        self._check_batch(x)
        ue = on_train if update_ema is None else update_ema
        check(lib.dssm_tower_forward(self._h, ptr(x.indptr), ptr(x.indices), ptr(x.values), int(on_train), int(ue), stream_ptr()))
        self._last = x  # keep the CSR alive for backward
        return self.tensor("loss")

    def eval_step(self, x: DeviceCSR, auc=None):
        """The reference's evaluation of one batch (new_dssm.py:274-286) from ONE forward: it runs sess.run(loss),
        sess.run(auc_op) and sess.run(auc_value) -- three full forwards with on_train=False; here the inference-mode forward
        runs once, the streaming AUC (a dssm_b200.DeviceStreamingAUC) is updated on the device from the same cos_sim_raw
        and nothing is copied to the host.  Returns (loss, auc) as device scalars (auc None without a metric)."""
        loss = self.forward(x, on_train=False)
        a = auc.update(self.tensor("cos_sim_raw"), self.conf.query_BS) if auc is not None else None
        return loss, a

    def backward(self) -> None:
        check(lib.dssm_tower_backward(self._h, stream_ptr()))

    def adam(self, grad_scale: float = 1.0) -> None:
        check(lib.dssm_tower_adam(self._h, float(grad_scale), stream_ptr()))

    # pipelined pieces for data-parallel training (dssm_b200/parallel.py)
    def backward_begin(self) -> None:
        """Backward without the dW1 gather: dense layers, small gradients, CSC build."""
        check(lib.dssm_tower_backward_begin(self._h, stream_ptr()))

    def backward_w1(self, chunk: int, n_chunks: int) -> None:
        """dW1 rows of column chunk `chunk` of `n_chunks` (a contiguous slice of the flat gradient buffer)."""
        check(lib.dssm_tower_backward_w1(self._h, chunk, n_chunks, stream_ptr()))

    def w1_chunk(self, chunk: int, n_chunks: int):
        off, cnt = C.c_int64(), C.c_int64()
        check(lib.dssm_tower_w1_chunk(self._h, chunk, n_chunks, C.byref(off), C.byref(cnt)))
        return off.value, cnt.value

    def adam_range(self, offset: int, count: int, grad_scale: float = 1.0) -> None:
        check(lib.dssm_tower_adam_range(self._h, offset, count, float(grad_scale), stream_ptr()))

    def adam_advance(self) -> None:
        check(lib.dssm_tower_adam_advance(self._h, stream_ptr()))

    def capture_graph_dp(self) -> None:
        """Capture training forward + backward_begin on the staging CSR into a CUDA graph."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            check(lib.dssm_tower_capture_graph_dp(self._h, s.cuda_stream))
        torch.cuda.current_stream(self.device).wait_stream(s)

    def fwd_bwd_begin_staged(self) -> torch.Tensor:
        check(lib.dssm_tower_fwd_bwd_begin_staged(self._h, stream_ptr()))
        return self.tensor("loss")

    def train_step(self, x: DeviceCSR) -> torch.Tensor:
        """sess.run(train_step, feed_dict=pull_batch(True, ...)) -- new_dssm.py:267-269.  Device CSR in, device loss out."""
        self._check_batch(x)
        check(lib.dssm_tower_train_step(self._h, ptr(x.indptr), ptr(x.indices), ptr(x.values), stream_ptr()))
        self._last = x
        return self.tensor("loss")

    # host-buffer path (what a caller holding scipy/numpy batches uses; bench.py's e2e)
    def pin(self, b: StackedBatch):
        """Pinned host copies of a stacked batch (for asynchronous upload)."""
        return (torch.from_numpy(b.indptr).pin_memory(), torch.from_numpy(b.indices).pin_memory(),
                torch.from_numpy(b.values).pin_memory(), b.nnz)

    def capture_graph(self) -> None:
        """Capture forward+backward+Adam on the staging CSR into a CUDA graph (replayed by train_step_host)."""
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            check(lib.dssm_tower_capture_graph(self._h, s.cuda_stream))
        torch.cuda.current_stream(self.device).wait_stream(s)
        self._graph = True

    def stage(self, x: DeviceCSR) -> None:
        """Device-to-device copy of a batch into the staging CSR (bench.py's device-resident `value` leg)."""
        self._check_batch(x)
        ip, ix, vl = self.staging_views()
        ip.copy_(x.indptr)
        ix[:x.nnz].copy_(x.indices)
        vl[:x.nnz].copy_(x.values)

    def staging_views(self):
        """torch views over the staging CSR inside the workspace (indptr, indices, values)."""
        ip, ix, vl = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib.dssm_tower_staging(self._h, C.byref(ip), C.byref(ix), C.byref(vl)))
        base = self.workspace.data_ptr()
        R = self.conf.rows
        w32 = self.workspace.view(torch.int32)
        o = lambda p: (p.value - base) // 4
        return (w32[o(ip):o(ip) + R + 1], w32[o(ix):o(ix) + self.max_nnz],
                self._ws_f32[o(vl):o(vl) + self.max_nnz])

    def train_step_staged(self) -> torch.Tensor:
        check(lib.dssm_tower_train_step_staged(self._h, stream_ptr()))
        return self.tensor("loss")

    def train_step_host(self, pinned, read_loss: bool = True) -> Optional[float]:
        """One step from HOST buffers: H2D of the CSR, the step, D2H of the loss (synchronises when read_loss)."""
        indptr, indices, values, nnz = pinned
        hl = ptr(self._host_loss) if read_loss else None
        check(lib.dssm_tower_train_step_host(self._h, ptr(indptr), ptr(indices), ptr(values), nnz, hl, stream_ptr()))
        return float(self._host_loss[0]) if read_loss else None

    # pipelined host feed: step k's upload overlaps step k-1's compute (include/dssm_b200.h)
    def train_step_host_async(self, pinned) -> int:
        """Issue one step from pinned HOST buffers without synchronising; returns its step id.  Keep at most two steps
        in flight: call feed_loss(step) (or feed_wait) on the one before last before issuing another."""
        indptr, indices, values, nnz = pinned
        if getattr(self, "_loss_slots", None) is None:
            self._loss_slots = torch.zeros(2, dtype=torch.float32).pin_memory()
        k_next = getattr(self, "_feed_next", 0)
        slot = C.c_void_p(self._loss_slots.data_ptr() + 4 * (k_next & 1))
        k = lib.dssm_tower_train_step_host_async(self._h, ptr(indptr), ptr(indices), ptr(values), nnz, slot, stream_ptr())
        if k < 0:
            raise DssmError(int(k), last_error())
        self._feed_next = k + 1
        return int(k)

    def feed_upload_async(self, pinned) -> int:
        """First half of train_step_host_async for callers that run their own step on the staging CSR (DataParallelTower)."""
        indptr, indices, values, nnz = pinned
        if getattr(self, "_loss_slots", None) is None:
            self._loss_slots = torch.zeros(2, dtype=torch.float32).pin_memory()
        k = lib.dssm_tower_feed_upload_async(self._h, ptr(indptr), ptr(indices), ptr(values), nnz, stream_ptr())
        if k < 0:
            raise DssmError(int(k), last_error())
        return int(k)

    def feed_step_done(self, step: int) -> None:
        slot = C.c_void_p(self._loss_slots.data_ptr() + 4 * (step & 1))
        check(lib.dssm_tower_feed_step_done(self._h, step, slot, stream_ptr()))

    def feed_wait(self, step: int) -> None:
        check(lib.dssm_tower_feed_wait(self._h, step))

    def feed_loss(self, step: int) -> float:
        """Loss of pipelined step `step` (blocks until that step is done)."""
        self.feed_wait(step)
        return float(self._loss_slots[step & 1])

    def train_epoch_host(self, pinned_batches) -> list:
        """The reference's inner training loop (new_dssm.py:261-269) over host batches with the double-buffered feed;
        returns every step's loss."""
        losses, prev = [], None
        for pb in pinned_batches:
            k = self.train_step_host_async(pb)
            if prev is not None:
                losses.append(self.feed_loss(prev))
            prev = k
        if prev is not None:
            losses.append(self.feed_loss(prev))
        return losses

    PHASES = ("spmm_fwd", "dense_fwd", "cos_loss", "dense_bwd", "csc_build", "dw_gather", "db1", "adam")

    def profile_step(self) -> Dict[str, float]:
        """One un-graphed step on the staging CSR with CUDA events between phases -> {phase: ms}. Synchronises."""
        buf = (C.c_float * 8)()
        check(lib.dssm_tower_profile_step(self._h, buf, stream_ptr()))
        return {k: float(buf[i]) for i, k in enumerate(self.PHASES)}

    def profile_step_overlapped(self) -> Dict[str, float]:
        """Same events around the step as it really runs (side stream, fused W1 Adam); 'csc_build' = wait at the join."""
        buf = (C.c_float * 8)()
        check(lib.dssm_tower_profile_step_overlapped(self._h, buf, stream_ptr()))
        return {k: float(buf[i]) for i, k in enumerate(self.PHASES)}

    def profile_timeline(self):
        """[(label, ms since the previous label)] for every call on the main stream of one real step, measured inside a
        CUDA graph with %globaltimer stamps (each stamp adds a ~1.5 us node).  Takes 3 training steps on the staging CSR."""
        names = C.create_string_buffer(4096)
        ms = (C.c_float * 128)()
        n = C.c_int32()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            check(lib.dssm_tower_profile_timeline(self._h, names, 4096, ms, 128, C.byref(n), s.cuda_stream))
        torch.cuda.current_stream(self.device).wait_stream(s)
        labels = names.value.decode().split(";")
        return [(labels[i], float(ms[i])) for i in range(n.value)]

    @property
    def launch_count(self) -> int:
        return int(lib.dssm_tower_launch_count(self._h))

    # ---- sess.run shim --------------------------------------------------------------------------------
    def run(self, fetches, feed_dict: Dict):
        """tower.run('train_step' | 'loss' | 'BN2/embedding_query_y:0' | [...], feed_dict=pull_batch(...))."""
        single = isinstance(fetches, str)
        names = [fetches] if single else list(fetches)
        x = self.to_device(stack_feed(feed_dict, self.conf))
        on_train = bool(feed_dict.get(ON_TRAIN, False))
        if any(n.split("/")[-1].split(":")[0] == "train_step" for n in names):
            self.train_step(x)
        else:
            self.forward(x, on_train=on_train)
        out = []
        for n in names:
            key = n.split("/")[-1].split(":")[0]
            if key == "train_step":
                out.append(None)
            elif key in ("auc_op", "update_op") or key in ("auc_value", "value"):
                # tf.metrics.auc's (value, update_op) pair of new_dssm.py:230: accumulated on the device, never reset
                if getattr(self, "auc", None) is None:
                    from .export import DeviceStreamingAUC

                    self.auc = DeviceStreamingAUC(self.device)
                v = self.auc.update(self.tensor("cos_sim_raw"), self.conf.query_BS) if key in ("auc_op", "update_op") else self.auc.result()
                out.append(np.float32(v.item()))
            else:
                out.append(self.tensor(n).detach().cpu().numpy().copy())
        return out[0] if single else out
