"""BoW documents -> CSR arrays (the HBM-resident corpus layout).

The reference keeps `documents` as list[list[(word_id, count)]] (stm.py:331-332) or any indexable
yielding such pairs (gensim MmCorpus, 03_fit_reference_model.py:73); its `create_dtm`
(stm.py:87-119) builds a scipy CSR from the same triples."""
import numpy as np


def pack_corpus(documents):
    """-> (doc_ptr int64 [D+1], word_id int32 [nnz], count float32 [nnz])"""
    if isinstance(documents, tuple) and len(documents) == 3:
        ptr, ids, cnt = documents
        return (np.ascontiguousarray(ptr, np.int64), np.ascontiguousarray(ids, np.int32),
                np.ascontiguousarray(cnt, np.float32))
    lens = np.fromiter((len(d) for d in documents), dtype=np.int64, count=len(documents))
    ptr = np.zeros(len(documents) + 1, dtype=np.int64)
    np.cumsum(lens, out=ptr[1:])
    nnz = int(ptr[-1])
    ids = np.empty(nnz, dtype=np.int32)
    cnt = np.empty(nnz, dtype=np.float32)
    pos = 0
    for doc in documents:
        n = len(doc)
        if n:
            arr = np.asarray(doc, dtype=np.float64).reshape(n, 2)
            ids[pos:pos + n] = arr[:, 0].astype(np.int64)
            cnt[pos:pos + n] = arr[:, 1]
            pos += n
    return ptr, ids, cnt


def word_counts(ptr, ids, cnt, V):
    """column sums of the document-term matrix (`STM.wcounts`, stm.py:485-486)"""
    return np.bincount(ids, weights=cnt.astype(np.float64), minlength=V)
