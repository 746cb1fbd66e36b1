"""The C-ABI library loads on a CPU-only box and exports every symbol include/dssm_b200.h declares
(no compute calls here -- those need a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "dssm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(dssm_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("dssm_spmm_fwd", "dssm_spmm_bwd_dw", "dssm_bn_forward", "dssm_fc_fwd", "dssm_merge_negative_doc",
                 "dssm_cos_softmax_loss", "dssm_adam_step", "dssm_corpus_topk", "dssm_topk_merge", "dssm_tower_train_step",
                 "dssm_tower_train_step_host"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from dssm_b200 import _lib

    assert _lib.LIB_PATH.exists()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in dssm_b200.h but not exported: {missing}"
    # and the Python binding covers the same set
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_version_and_error_string():
    from dssm_b200 import _lib

    assert _lib.lib.dssm_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_config_struct_layout_matches_header():
    from dssm_b200 import _lib

    # 1 + 1 + 8 + 6 int32 + 8 floats
    assert ctypes.sizeof(_lib.dssm_config) == 4 * (2 + 8 + 6 + 8)


def test_bad_arguments_fail_loudly_without_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on the CPU box."""
    from dssm_b200 import _lib

    rc = _lib.lib.dssm_spmm_fwd(None, None, None, 4, 8, None, None, 4, None, None)
    assert rc == 1 and "null" in _lib.last_error()
    cfg = _lib.dssm_config()
    h = ctypes.c_void_p()
    assert _lib.lib.dssm_tower_create(ctypes.byref(cfg), ctypes.byref(h)) == 1  # n_layers = 0
    with pytest.raises(_lib.DssmError):
        _lib.check(3)


def test_tower_layout_queries_work_without_gpu():
    """create / layout / workspace sizing are host-only."""
    from dssm_b200 import Config, _lib
    from dssm_b200.tower import _make_c_config

    conf = Config(TRIGRAM_D=1000, query_BS=8, NEG=3, layers=(300, 300, 128))
    cc = _make_c_config(conf)
    h = ctypes.c_void_p()
    _lib.check(_lib.lib.dssm_tower_create(ctypes.byref(cc), ctypes.byref(h)))
    try:
        P = _lib.lib.dssm_tower_param_count(h)
        want = 1000 * 300 + 300 + 300 * 300 + 300 + 300 * 128 + 128 + 2 * 2 * (300 + 300 + 128)
        assert P >= want and P - want < 4 * 20
        assert _lib.lib.dssm_tower_ema_count(h) == 2 * 2 * (300 + 300 + 128)
        assert _lib.lib.dssm_tower_workspace_bytes(h, 4096) > 0
        names = []
        buf = ctypes.create_string_buffer(64)
        for i in range(_lib.lib.dssm_tower_num_tensors(h, 0)):
            _lib.check(_lib.lib.dssm_tower_tensor_info(h, 0, i, buf, 64, None, None, None))
            names.append(buf.value.decode())
        assert names[:4] == ["W1", "b1", "W2", "b2"] and "bn3_beta" in names
    finally:
        _lib.lib.dssm_tower_destroy(h)
