"""Text -> bag-of-grams CSR, compatible with the vectorizer the reference fits (new_dssm.py:33-44):

    vectorizer = CountVectorizer(token_pattern=r"(?u)\\b\\w+\\b")
    vectorizer.fit(ad_act + bhv_act + ad_act_neg)          # vocabulary = sorted set of tokens
    query_train_dat = vectorizer.transform(bhv_act)        # scipy CSR, int64 term counts, sorted column indices
    TRIGRAM_D = len(vectorizer.get_feature_names())

scikit-learn is a dependency of the reference's *script*, not of the hot path; this module restates the four defaults
that decide the matrix (lowercase=True, analyzer='word', ngram_range=(1, 1), binary=False; vocabulary in sorted token
order; out-of-vocabulary tokens dropped at transform time) so that the loader needs nothing but NumPy/SciPy, and
tests/test_vectorizer.py checks vocabulary and matrices against sklearn itself (present in this image) token for token.
`max_features` (dssm_no_bn/my_dssm.py:37) keeps the most frequent terms; sklearn breaks frequency ties with an unstable
argsort, here ties go to the alphabetically first term.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import scipy.sparse as sp

TOKEN_PATTERN = r"(?u)\b\w+\b"  # the reference's (new_dssm.py:33); sklearn's default needs two characters per token


class CountVectorizerCompat:
    def __init__(self, token_pattern: str = TOKEN_PATTERN, lowercase: bool = True, max_features: Optional[int] = None,
                 vocabulary: Optional[Dict[str, int]] = None, dtype=np.int64):
        self.token_pattern = token_pattern
        self.lowercase = lowercase
        self.max_features = max_features
        self.dtype = dtype
        self._tok = re.compile(token_pattern)
        self.vocabulary_: Optional[Dict[str, int]] = dict(vocabulary) if vocabulary is not None else None

    # ---- analysis ---------------------------------------------------------------------------------------------
    def tokenize(self, doc: str) -> List[str]:
        return self._tok.findall(doc.lower() if self.lowercase else doc)

    # ---- fit --------------------------------------------------------------------------------------------------
    def fit(self, raw_documents: Iterable[str]) -> "CountVectorizerCompat":
        counts: Dict[str, int] = {}
        for doc in raw_documents:
            for t in self.tokenize(doc):
                counts[t] = counts.get(t, 0) + 1
        terms = sorted(counts)
        if self.max_features is not None and len(terms) > self.max_features:
            # most frequent first (total term frequency, as sklearn's _limit_features); ties: alphabetical
            order = sorted(terms, key=lambda t: (-counts[t], t))[: self.max_features]
            terms = sorted(order)
        if not terms:
            raise ValueError("empty vocabulary; perhaps the documents only contain stop words")
        self.vocabulary_ = {t: i for i, t in enumerate(terms)}
        return self

    def get_feature_names_out(self) -> np.ndarray:
        return np.asarray(sorted(self.vocabulary_, key=self.vocabulary_.get), dtype=object)

    def get_feature_names(self) -> List[str]:  # the (removed) sklearn method new_dssm.py:44 calls
        return list(self.get_feature_names_out())

    # ---- transform ----------------------------------------------------------------------------------------------
    def transform(self, raw_documents: Sequence[str]) -> sp.csr_matrix:
        if self.vocabulary_ is None:
            raise ValueError("vocabulary not fitted")
        vocab = self.vocabulary_
        indptr = [0]
        cols: List[int] = []
        for doc in raw_documents:
            for t in self.tokenize(doc):
                j = vocab.get(t)
                if j is not None:
                    cols.append(j)
            indptr.append(len(cols))
        n_docs = len(indptr) - 1
        m = sp.csr_matrix((np.ones(len(cols), dtype=self.dtype), np.asarray(cols, dtype=np.int32), np.asarray(indptr, dtype=np.int32)),
                          shape=(n_docs, len(vocab)))
        m.sum_duplicates()  # term counts; also sorts the column indices of every row
        m.sort_indices()
        return m

    def fit_transform(self, raw_documents: Sequence[str]) -> sp.csr_matrix:
        docs = list(raw_documents)
        return self.fit(docs).transform(docs)


def char_split(text: str) -> str:
    """The reference feeds character-split text ("a b c"): get_data_set_comment joins the characters of the cleaned
    string with spaces (utils/utils.py:395-398) before the vectorizer sees it."""
    return " ".join(text)
