"""Host-side callers of the hot path: streaming AUC (tf.metrics.auc semantics), accuracy, the idx:val dump format."""
import numpy as np

from dssm_b200.export import StreamingAUC, accuracy, format_mid_vector, labels_for, write_mid_vectors


def exact_auc(labels, pred):
    """Mann-Whitney AUC (ties count one half)."""
    labels = np.asarray(labels).astype(bool)
    p, n = pred[labels], pred[~labels]
    return float(((p[:, None] > n[None, :]).sum() + 0.5 * (p[:, None] == n[None, :]).sum()) / (p.size * n.size))


def test_streaming_auc_close_to_exact_and_accumulates():
    rng = np.random.default_rng(0)
    B, NEG = 100, 4
    lab = labels_for(B, NEG)
    pred1 = np.clip(rng.normal(0.5 + 0.15 * lab, 0.15), 0, 1).astype(np.float32)
    m = StreamingAUC(2000)
    a1 = m.update(lab, pred1)
    assert abs(a1 - exact_auc(lab, pred1)) < 2e-3  # 2000-threshold trapezoid vs exact
    pred2 = np.clip(rng.normal(0.5 - 0.05 * lab, 0.2), 0, 1).astype(np.float32)
    a2 = m.update(lab, pred2)  # never reset: this is the AUC over both batches, as in the reference's epoch loop
    both = exact_auc(np.concatenate([lab, lab]), np.concatenate([pred1, pred2]))
    assert abs(a2 - both) < 2e-3
    assert m.tp[0] == 2 * B and m.fp[0] == 2 * B * NEG  # threshold -1e-7: everything positive
    assert m.tp[-1] == 0 and m.fp[-1] == 0  # threshold 1+1e-7: nothing positive


def test_auc_thresholds_match_tf():
    m = StreamingAUC(5)
    np.testing.assert_allclose(m.thresholds, [-1e-7, 0.25, 0.5, 0.75, 1 + 1e-7], rtol=0, atol=1e-7)  # float32, like TF


def test_accuracy_and_labels():
    prob = np.array([[0.6, 0.4], [0.2, 0.8], [0.5, 0.5]])
    assert abs(accuracy(prob) - 2 / 3) < 1e-12  # argmax takes the first maximum, like tf.argmax
    assert labels_for(2, 3).tolist() == [1, 1, 0, 0, 0, 0, 0, 0]


def test_mid_vector_format(tmp_path):
    v = np.array([0.0, 0.123456789, 5e-5, 1.5, 0.00011], dtype=np.float32)
    s = format_mid_vector(v)
    # value strings are Python's str(float) of the float32 value widened to double, cut to 6 characters
    assert s == "1:" + str(float(v[1]))[:6] + ",3:1.5,4:" + str(float(v[4]))[:6]
    p = tmp_path / "y_mid_vector.txt"
    write_mid_vectors(str(p), ["a b c", "d"], np.stack([v, np.zeros(5, np.float32)]))
    write_mid_vectors(str(p), ["e"], v[None])  # 'a+' mode appends
    lines = p.read_text().splitlines()
    assert lines[0] == "abc\t" + s and lines[1] == "d\t" and lines[2] == "e\t" + s
