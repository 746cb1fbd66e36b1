"""Accuracy / Auc stages on the device (csrc/metrics.cu; new_dssm.py:219-231) against the host restatement of
tf.metrics.auc (dssm_b200.export.StreamingAUC), the one-forward evaluation step and the sess.run shim's auc fetches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_auc_equals_host_streaming_auc():
    from dssm_b200.export import DeviceStreamingAUC, StreamingAUC, labels_for

    rng = np.random.default_rng(0)
    host, dev = StreamingAUC(2000), DeviceStreamingAUC("cuda", 2000)
    B, NEG = 257, 5
    for step in range(4):  # never reset: the counters accumulate over updates (new_dssm.py:252)
        pos = np.clip(rng.normal(0.6, 0.25, B), -0.2, 1.2)
        neg = np.clip(rng.normal(0.3, 0.25, B * NEG), -0.2, 1.2)
        pred = np.concatenate([pos, neg]).astype(np.float32)
        pred[rng.integers(0, pred.size, 6)] = np.nan  # 0/0 cosines of all-zero relu embeddings
        thr = host.thresholds
        pred[rng.integers(0, pred.size, 40)] = thr[rng.integers(0, thr.size, 40)]  # exactly on a threshold: `>` not `>=`
        if step == 2:
            pred[:10] = [0.0, 1.0, -1e-7, 1.0 + 1e-7, 2.0, -1.0, np.inf, -np.inf, 0.5, np.float32(1 / 1999)]
        want = host.update(labels_for(B, NEG), pred)
        got = dev.update(torch.from_numpy(pred).cuda(), B).item()
        assert abs(got - want) <= 1e-12, (step, got, want)
        # the four confusion counters themselves, from the two histograms
        ph, nh = dev.pos_hist.cpu().numpy(), dev.neg_hist.cpu().numpy()
        tp = ph[::-1].cumsum()[::-1][1:]  # entries with bucket > i
        fp = nh[::-1].cumsum()[::-1][1:]
        assert np.array_equal(tp, host.tp.astype(np.int64)) and np.array_equal(fp, host.fp.astype(np.int64))
        assert np.array_equal(ph.sum() - tp, host.fn.astype(np.int64)) and np.array_equal(nh.sum() - fp, host.tn.astype(np.int64))
    assert abs(dev.result().item() - host.result()) <= 1e-12


def test_eval_step_is_one_forward_and_matches_three_reference_runs():
    """new_dssm.py:274-286 evaluates a batch with three sess.run calls (loss, auc_op, auc_value), each a full forward.
    eval_step gives the same loss and the same running AUC from one inference-mode forward, on the device."""
    from dssm_b200 import Config, DSSMTower, pull_batch
    from dssm_b200.export import DeviceStreamingAUC, StreamingAUC, labels_for
    from dssm_b200.synthetic import init_params, make_batch

    conf = Config(TRIGRAM_D=5000, query_BS=64, NEG=4, layers=(64, 32))
    params = init_params(conf, 0)
    batches = [make_batch(conf, s, 8, 16) for s in range(3)]
    mx = max(b.nnz for b in batches)
    t = DSSMTower(conf, max_nnz=mx, params=params)
    for b in batches[:2]:
        t.train_step(t.to_device(b))  # move the weights and fill the EMA shadows
    auc = DeviceStreamingAUC(t.device)
    host = StreamingAUC()
    u = DSSMTower(conf, max_nnz=mx, params=params)
    u.load_state_dict(t.state_dict())
    for b in batches:
        n0 = t.launch_count
        loss, a = t.eval_step(t.to_device(b), auc)
        fwd_launches = t.launch_count - n0
        # the reference way through the sess.run shim: three runs, three forwards
        X = b.to_scipy()
        B = conf.query_BS
        feed = pull_batch(False, X[:B], X[B:2 * B], X[2 * B:], 0, B, conf=conf)
        n1 = u.launch_count
        loss_v = u.run("loss", feed)
        u.run("Auc/auc/update_op:0", feed)
        auc_v = u.run("Auc/auc/value:0", feed)
        assert u.launch_count - n1 == 3 * fwd_launches
        assert loss.item() == float(np.asarray(loss_v).reshape(-1)[0])
        want = host.update(labels_for(B, conf.NEG), t.tensor("cos_sim_raw").cpu().numpy())
        assert abs(a.item() - want) <= 1e-12 and abs(float(auc_v) - want) <= 1e-6
    acc = t.tensor("accuracy").item()
    prob = t.tensor("prob").cpu().numpy()
    assert abs(acc - float(np.mean(np.argmax(prob, axis=1) == 0))) < 1e-7
