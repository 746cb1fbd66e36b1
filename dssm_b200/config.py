"""Hyper-parameters of the DSSM tower.  Attribute names are the reference's
(semantic_matching/dssm/config.py:19-28: query_BS, L1_N, L2_N, learning_rate, NEG; new_dssm.py:44:
TRIGRAM_D); switches cover the variants SURVEY.md section 8a lists (dssm_no_bn, loss epsilon,
un-normalised loss, tanh, a third layer)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple


class Config(object):
    def to_string(self):
        print("conf params: ")
        hyper_params = self.__dict__
        for key in hyper_params:
            print(str(key) + ": " + str(hyper_params[key]))

    def __init__(self, TRIGRAM_D: int = 0, query_BS: int = 400, L1_N: int = 100, L2_N: int = 100, NEG: int = 4,
                 learning_rate: float = 0.01, layers: Optional[Sequence[int]] = None, use_bn: bool = True,
                 act: str = "relu", loss_eps: float = 0.0, loss_div_bs: bool = True, gemm_mode: str = "fp32",
                 verbose: bool = False):
        self.TRIGRAM_D = int(TRIGRAM_D)
        self.query_BS = int(query_BS)  # config.py:19
        self.L1_N = int(L1_N)  # config.py:20
        self.L2_N = int(L2_N)  # config.py:21
        self.learning_rate = float(learning_rate)  # config.py:23
        self.NEG = int(NEG)  # config.py:28
        # the reference graph has exactly (L1_N, L2_N); "300-300-128" needs a layer list
        self.layers: Tuple[int, ...] = tuple(int(x) for x in layers) if layers is not None else (self.L1_N, self.L2_N)
        if layers is not None and len(self.layers) >= 1:
            self.L1_N = self.layers[0]
            if len(self.layers) >= 2:
                self.L2_N = self.layers[1]
        self.use_bn = bool(use_bn)  # False: semantic_matching/dssm_no_bn/my_dssm.py:98-121
        self.act = act  # reference: relu
        self.bn_eps = 1e-3  # new_dssm.py:87
        self.ema_decay = 0.5  # new_dssm.py:78
        self.gamma = 20.0  # new_dssm.py:199
        self.loss_eps = float(loss_eps)  # 1e-8: dssm_no_bn/my_dssm.py:169
        self.loss_div_bs = bool(loss_div_bs)  # False: archive/dssm_v2.py:184
        self.beta1, self.beta2, self.adam_eps = 0.9, 0.999, 1e-8  # tf.train.AdamOptimizer defaults
        self.gemm_mode = gemm_mode
        if verbose:
            self.to_string()

    @property
    def rows(self) -> int:
        return (2 + self.NEG) * self.query_BS

    def layer_dims(self):
        dims, d_in = [], self.TRIGRAM_D
        for n in self.layers:
            dims.append((d_in, n))
            d_in = n
        return dims


# BASELINE.json configs (C1..C4)
def baseline_config(name: str, gemm_mode: str = "tc_3xtf32") -> Config:
    c = _baseline_config(name)
    c.gemm_mode = gemm_mode
    return c


def _baseline_config(name: str) -> Config:
    name = name.upper()
    if name == "C1":
        return Config(TRIGRAM_D=21128, query_BS=100, NEG=4, layers=(300, 300, 128))
    if name == "C2":
        return Config(TRIGRAM_D=49284, query_BS=1024, NEG=4, layers=(300, 300, 128))
    if name == "C3":
        return Config(TRIGRAM_D=49284, query_BS=8192, NEG=4, layers=(300, 300, 128))
    if name == "C4":
        return Config(TRIGRAM_D=49284, query_BS=1024, NEG=50, layers=(300, 300, 128))
    if name == "C4_NOBN":
        return Config(TRIGRAM_D=49284, query_BS=1024, NEG=50, layers=(300, 300, 128), use_bn=False, loss_eps=1e-8)
    raise KeyError(name)
