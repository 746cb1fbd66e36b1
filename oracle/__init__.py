"""CPU oracle for the DSSM two-tower hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  ``dssm_b200`` never imports this package.

PARITY UNPINNED: the reference (MC-Zealot/dssm) ships no tests, no golden
vectors and no fixtures, and its arithmetic lives in an un-vendored, un-pinned
TensorFlow 1.x wheel that is not installable here.  The
oracle is therefore a restatement of the reference *graph* (file:line cited on
every function) under documented TF-1.x semantics, cross-checked against an
independent torch-CPU float64 autograd model and a literal replay of the
reference's Python loops, not against outputs of the reference itself.
"""
from .dssm_oracle import OracleConfig, DSSMOracle, init_params, DPOracle  # noqa: F401
from .literal_replay import (  # noqa: F401
    merge_negative_doc_literal,
    cosine_similarity_literal,
    loss_literal,
)
from .retrieval_oracle import exact_cosine_scores, corpus_topk_oracle, merge_topk_oracle  # noqa: F401
