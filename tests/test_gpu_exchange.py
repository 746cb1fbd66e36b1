"""The NVLink pull / Adam / push kernel (csrc/nvlink.cu: dssm_w1_shard_reduce_adam) on ONE GPU: the kernel takes pointer
tables, so n replicas are emulated by n buffers on the same device -- every "rank" runs its owner pass over its row
shard, in turn, exactly as the ranks of a data-parallel step do between the two barriers.  Checked against TF-Adam on the
rank-ordered mean gradient; all replicas of W1 must end BIT-identical."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_ranks", [2, 4, 8])
def test_shard_reduce_adam_emulated_ranks(n_ranks):
    from dssm_b200._lib import check, lib, ptr, stream_ptr
    from tests.helpers import tf_adam_reference

    D, L1 = 1003, 300
    g = torch.Generator(device="cuda").manual_seed(n_ranks)
    W0 = torch.randn((D, L1), generator=g, device="cuda") * 0.05
    Ws = [W0.clone() for _ in range(n_ranks)]
    dWs = [torch.randn((D, L1), generator=g, device="cuda") * 1e-3 for _ in range(n_ranks)]
    m0 = torch.randn((D, L1), generator=g, device="cuda") * 1e-4
    v0 = torch.rand((D, L1), generator=g, device="cuda") * 1e-6
    ms, vs = [m0.clone() for _ in range(n_ranks)], [v0.clone() for _ in range(n_ranks)]
    beta_pow = torch.tensor([0.9 ** 3, 0.999 ** 3], dtype=torch.float32, device="cuda")
    arr = C.c_void_p * n_ranks
    peer_dw, peer_w = arr(*[d.data_ptr() for d in dWs]), arr(*[w.data_ptr() for w in Ws])
    per = (D + n_ranks - 1) // n_ranks
    for r in range(n_ranks):  # rank r owns rows [r*per, (r+1)*per) and keeps m, v only for them
        lo, hi = min(r * per, D), min((r + 1) * per, D)
        check(lib.dssm_w1_shard_reduce_adam(peer_dw, peer_w, n_ranks, r, D, L1, lo, hi, ptr(ms[r]), ptr(vs[r]), ptr(beta_pow),
                                            0.01, 0.9, 0.999, 1e-8, stream_ptr()))
    torch.cuda.synchronize()
    for r in range(1, n_ranks):
        assert torch.equal(Ws[r], Ws[0]), f"replica {r} of W1 differs from replica 0"
    # reference: rank-ordered fp32 sum, then the TF-Adam formula with grad_scale 1/n
    gsum = dWs[0].clone()
    for r in range(1, n_ranks):
        gsum += dWs[r]
    bp = beta_pow.cpu().numpy()
    rp, rm, rv = tf_adam_reference(W0.cpu().numpy(), gsum.cpu().numpy(), m0.cpu().numpy(), v0.cpu().numpy(), bp[0], bp[1],
                                   grad_scale=1.0 / n_ranks)
    got_m = torch.cat([ms[r][min(r * per, D):min((r + 1) * per, D)] for r in range(n_ranks)]).cpu().numpy()
    got_v = torch.cat([vs[r][min(r * per, D):min((r + 1) * per, D)] for r in range(n_ranks)]).cpu().numpy()
    assert np.abs(got_m - rm).max() <= 1e-6 * np.abs(rm).max()
    assert np.abs(got_v - rv).max() <= 1e-6 * np.abs(rv).max()
    assert np.abs(Ws[0].cpu().numpy() - rp).max() <= 2e-5 * 0.01 + 1e-6 * np.abs(rp).max()
    # rows outside a rank's shard keep their optimizer state untouched
    for r in range(n_ranks):
        lo, hi = min(r * per, D), min((r + 1) * per, D)
        mask = torch.ones(D, dtype=torch.bool, device="cuda")
        mask[lo:hi] = False
        assert torch.equal(ms[r][mask], m0[mask]) and torch.equal(vs[r][mask], v0[mask])
