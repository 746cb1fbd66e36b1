"""ctypes binding of libdssm_b200.so (include/dssm_b200.h).

There is no CPU fallback: if the library is missing and cannot be built with nvcc, importing this module
raises.  PyTorch tensors are used only as device buffers -- every call passes raw pointers and sizes.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libdssm_b200.so"

DSSM_OK = 0
ACT = {"none": 0, None: 0, "relu": 1, "tanh": 2}
GEMM = {"fp32": 0, "tc_3xtf32": 1, "tc_tf32": 2}
MAX_LAYERS = 8


class DssmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dssm_b200 error {code}: {msg}")
        self.code = code


class dssm_config(C.Structure):
    _fields_ = [
        ("TRIGRAM_D", C.c_int32),
        ("n_layers", C.c_int32),
        ("layers", C.c_int32 * MAX_LAYERS),
        ("NEG", C.c_int32),
        ("query_BS", C.c_int32),
        ("use_bn", C.c_int32),
        ("act", C.c_int32),
        ("loss_div_bs", C.c_int32),
        ("gemm_mode", C.c_int32),
        ("bn_eps", C.c_float),
        ("ema_decay", C.c_float),
        ("gamma", C.c_float),
        ("loss_eps", C.c_float),
        ("learning_rate", C.c_float),
        ("beta1", C.c_float),
        ("beta2", C.c_float),
        ("adam_eps", C.c_float),
    ]


def _load() -> C.CDLL:
    from . import build as _build

    if os.environ.get("DSSM_B200_REBUILD") == "1" or not _build.up_to_date():
        _build.build(force=os.environ.get("DSSM_B200_REBUILD") == "1")  # needs nvcc; raises if absent
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} is missing and could not be built; dssm_b200 has no CPU fallback")
    return C.CDLL(str(LIB_PATH))


_p, _i32, _i64, _f, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); mirrors include/dssm_b200.h one to one
SIGNATURES = {
    "dssm_last_error": (C.c_char_p, []),
    "dssm_version": (C.c_int, []),
    "dssm_sm_count": (C.c_int, []),
    "dssm_spmm_fwd": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _p, _i32, _p, _p]),
    "dssm_spmm_bwd_dw_workspace_bytes": (_sz, [_i32, _i32, _i32, _i64]),
    "dssm_spmm_bwd_dw": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _i32, _p, _i32, _p, _sz, _p]),
    "dssm_spmm_bwd_csc_build": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p, _sz, _p]),
    "dssm_spmm_bwd_dw_range": (C.c_int, [_p, _i32, _i32, _i32, _p, _i32, _i32, _i32, _p, _sz, _p]),
    "dssm_spmm_bwd_dw_adam": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, _p, _f, _f, _f, _f, _i32, _p, _sz, _p]),
    "dssm_spmm_bwd_adam_absent": (C.c_int, [_i32, _i32, _p, _p, _p, _p, _f, _f, _f, _f, _p, _sz, _p]),
    "dssm_w1_shard_reduce_adam": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _f, _f, _f, _f, _p]),
    "dssm_w1_shard_reduce_adam_mc": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _f, _f, _f, _f, _p]),
    "dssm_w1_slots_reduce_adam": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _f, _f, _f, _f, _p]),
    "dssm_tower_set_w1_push": (C.c_int, [_p, _i32]),
    "dssm_tower_backward_w1_push": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _i32, _p]),
    "dssm_peer_flags_bytes": (_sz, []),
    "dssm_peer_signal": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p]),
    "dssm_peer_wait": (C.c_int, [_p, _i32, _i32, _i32, _p]),
    "dssm_peer_epoch_advance": (C.c_int, [_p, _p]),
    "dssm_peer_barrier": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p]),
    "dssm_bn_workspace_bytes": (_sz, [_i32, _i32]),
    "dssm_bn_forward": (C.c_int, [_p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "dssm_bn_act_apply": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _i32, _p, _p]),
    "dssm_bn_act_backward": (C.c_int, [_p, _p, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "dssm_fc_fwd_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "dssm_fc_fwd": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _i32, _p, _p, _i32, _p, _i32, _p, _sz, _p]),
    "dssm_fc_bwd_dx": (C.c_int, [_p, _i32, _i32, _p, _i32, _p, _i32, _p, _sz, _p]),
    "dssm_fc_bwd_dw_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "dssm_fc_bwd_dw": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _i32, _p, _i32, _p, _p, _i32, _p, _sz, _p]),
    "dssm_colsum_workspace_bytes": (_sz, [_i32, _i32]),
    "dssm_colsum": (C.c_int, [_p, _i32, _i32, _p, _p, _sz, _p]),
    "dssm_merge_negative_doc": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "dssm_merge_negative_doc_index": (C.c_int, [_i32, _i32, _p, _p]),
    "dssm_cos_softmax_loss": (C.c_int, [_p, _i32, _i32, _i32, _f, _f, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "dssm_cos_softmax_loss_fused": (C.c_int, [_p, _p, _p, _i32, _p, _i32, _i32, _i32, _f, _f, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "dssm_adam_step": (C.c_int, [_p, _p, _p, _p, _i64, _p, _f, _f, _f, _f, _f, _p]),
    "dssm_adam_advance": (C.c_int, [_p, _f, _f, _p]),
    "dssm_auc_update": (C.c_int, [_p, _i32, _i32, _p, _i32, _p, _p, _p]),
    "dssm_auc_result": (C.c_int, [_p, _p, _i32, _p, _p]),
    "dssm_accuracy": (C.c_int, [_p, _i32, _i32, _p, _p]),
    "dssm_corpus_topk_workspace_bytes": (_sz, [_i32, _i64, _i32, _i32]),
    "dssm_corpus_topk": (C.c_int, [_p, _i32, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _sz, _p]),
    "dssm_corpus_topk_tc_workspace_bytes": (_sz, [_i32, _i64, _i32, _i32]),
    "dssm_corpus_topk_tc": (C.c_int, [_p, _i32, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _sz, _p]),
    "dssm_corpus_index_bytes": (_sz, [_i64, _i32]),
    "dssm_corpus_index_build": (C.c_int, [_p, _i64, _i32, _p, _sz, _p]),
    "dssm_corpus_topk_indexed_workspace_bytes": (_sz, [_i32, _i32]),
    "dssm_corpus_topk_indexed": (C.c_int, [_p, _i32, _p, _p, _i64, _i32, _i32, _i32, _p, _p, _p, _p, _sz, _p]),
    "dssm_topk_merge": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p]),
    "dssm_tower_create": (C.c_int, [C.POINTER(dssm_config), C.POINTER(_p)]),
    "dssm_tower_destroy": (None, [_p]),
    "dssm_tower_param_count": (_i64, [_p]),
    "dssm_tower_ema_count": (_i64, [_p]),
    "dssm_tower_workspace_bytes": (_sz, [_p, _i64]),
    "dssm_tower_num_tensors": (_i32, [_p, _i32]),
    "dssm_tower_tensor_info": (C.c_int, [_p, _i32, _i32, C.c_char_p, _i32, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "dssm_tower_bind": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _sz, _i64]),
    "dssm_tower_syncbn_bytes": (_sz, [_p, _i32]),
    "dssm_tower_set_syncbn": (C.c_int, [_p, _i32, _i32, _p]),
    "dssm_tower_forward": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _p]),
    "dssm_tower_backward": (C.c_int, [_p, _p]),
    "dssm_tower_adam": (C.c_int, [_p, _f, _p]),
    "dssm_tower_backward_begin": (C.c_int, [_p, _p]),
    "dssm_tower_backward_w1": (C.c_int, [_p, _i32, _i32, _p]),
    "dssm_tower_w1_chunk": (C.c_int, [_p, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64)]),
    "dssm_tower_adam_range": (C.c_int, [_p, _i64, _i64, _f, _p]),
    "dssm_tower_adam_advance": (C.c_int, [_p, _p]),
    "dssm_tower_capture_graph_dp": (C.c_int, [_p, _p]),
    "dssm_tower_fwd_bwd_begin_staged": (C.c_int, [_p, _p]),
    "dssm_tower_train_step": (C.c_int, [_p, _p, _p, _p, _p]),
    "dssm_tower_train_step_host": (C.c_int, [_p, _p, _p, _p, _i64, _p, _p]),
    "dssm_tower_capture_graph": (C.c_int, [_p, _p]),
    "dssm_tower_staging": (C.c_int, [_p, C.POINTER(_p), C.POINTER(_p), C.POINTER(_p)]),
    "dssm_tower_train_step_staged": (C.c_int, [_p, _p]),
    "dssm_tower_train_step_host_async": (C.c_int64, [_p, _p, _p, _p, _i64, _p, _p]),
    "dssm_tower_feed_wait": (C.c_int, [_p, _i64]),
    "dssm_tower_feed_upload_async": (C.c_int64, [_p, _p, _p, _p, _i64, _p]),
    "dssm_tower_feed_step_done": (C.c_int, [_p, _i64, _p, _p]),
    "dssm_host_stack_csr": (_i64, [_i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32]),
    "dssm_tower_launch_count": (_i64, [_p]),
    "dssm_tower_profile_step": (C.c_int, [_p, C.POINTER(C.c_float), _p]),
    "dssm_tower_profile_step_overlapped": (C.c_int, [_p, C.POINTER(C.c_float), _p]),
    "dssm_tower_profile_timeline": (C.c_int, [_p, C.c_char_p, _i32, C.POINTER(C.c_float), _i32, C.POINTER(_i32), _p]),
}

class _LazyLib:
    """libdssm_b200.so, mapped on the first attribute access (not at import): modules that only need the host-side
    helpers (config, synthetic batches -- e.g. bench.py's --impl reference arm) never load the product library.  The
    first access loads it, binds every SIGNATURES entry (AttributeError = header and library out of sync) and then
    serves attributes straight from the CDLL.  There is still no CPU fallback: a missing library raises here."""

    _cdll = None

    def _bind(self) -> C.CDLL:
        if _LazyLib._cdll is None:
            cdll = _load()
            for _name, (_res, _args) in SIGNATURES.items():
                _fn = getattr(cdll, _name)
                _fn.restype = _res
                _fn.argtypes = _args
            _LazyLib._cdll = cdll
        return _LazyLib._cdll

    def __getattr__(self, name):
        fn = getattr(self._bind(), name)
        self.__dict__[name] = fn
        return fn


lib = _LazyLib()


def loaded() -> bool:
    """True once the shared library has been mapped into this process."""
    return _LazyLib._cdll is not None


def last_error() -> str:
    return lib.dssm_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != DSSM_OK:
        raise DssmError(rc, last_error())


def ptr(t) -> int:
    """Device (or pinned host) pointer of a torch tensor / None."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(stream=None) -> int:
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream
