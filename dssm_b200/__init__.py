"""dssm_b200 -- B200-native DSSM two-tower hot path behind the reference's model-building surface.

The first call into the C ABI loads libdssm_b200.so (building it with nvcc if absent); there is no CPU fallback.
"""
from ._lib import DssmError, LIB_PATH, lib  # noqa: F401
from .config import Config, baseline_config  # noqa: F401
from .batch import (  # noqa: F401
    SparseTensorValue, StackedBatch, bow_csr, convert_seq2bow, convert_sparse_matrix_to_sparse_tensor, load_vocab,
    pull_batch, stack_csr, stack_feed,
)
from .ops import (  # noqa: F401
    BNState, Cosine_Similarity, DeviceCSR, Loss, Merge_Negative_Doc, add_layer, batch_normalization,
    merge_negative_doc_index,
)
from .tower import DSSMTower  # noqa: F401
from .parallel import DataParallelTower, shard_stacked_batch  # noqa: F401
from .retrieval import CorpusIndex, corpus_topk, sharded_corpus_topk, topk_merge  # noqa: F401
from .export import DeviceStreamingAUC, StreamingAUC, embed_docs, embed_queries, write_mid_vectors  # noqa: F401
from .loader import HostBatchLoader  # noqa: F401
from .vectorizer import CountVectorizerCompat, char_split  # noqa: F401

__version__ = "0.1.0"
