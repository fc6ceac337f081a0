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


def read_mm(path):
    """A gensim `MmCorpus` file (Matrix Market coordinate format, documents x words, 1-based; written by
    `corpora.MmCorpus.serialize`, 02_create_corpus.py:42 — e.g. the reference's artifacts/wiki_data/BoW_corpus.mm,
    read back at 03_fit_reference_model.py:43-46) -> (doc_ptr, word_id, count, V).  Duplicate (document, word)
    entries are summed, ids come out ascending within a document."""
    D = V = nnz = None
    rows, cols, vals = [], [], []
    with open(path) as f:
        header = f.readline()
        if not header.startswith("%%MatrixMarket") or "coordinate" not in header:
            raise ValueError("not a Matrix Market coordinate file")
        for line in f:
            if line.startswith("%") or not line.strip():
                continue
            parts = line.split()
            if D is None:
                D, V, nnz = int(parts[0]), int(parts[1]), int(parts[2])
                continue
            rows.append(int(parts[0]) - 1)
            cols.append(int(parts[1]) - 1)
            vals.append(float(parts[2]))
    if D is None:
        raise ValueError("missing size line")
    rows, cols, vals = np.asarray(rows, np.int64), np.asarray(cols, np.int64), np.asarray(vals, np.float64)
    if len(rows) != nnz:
        raise ValueError(f"expected {nnz} entries, found {len(rows)}")
    if len(rows) and (rows.min() < 0 or rows.max() >= D or cols.min() < 0 or cols.max() >= V):
        raise ValueError("index out of range")
    key = rows * V + cols
    uk, inv = np.unique(key, return_inverse=True)
    cnt = np.bincount(inv, weights=vals, minlength=len(uk))
    ptr = np.zeros(D + 1, dtype=np.int64)
    np.cumsum(np.bincount(uk // V, minlength=D), out=ptr[1:])
    return ptr, (uk % V).astype(np.int32), cnt.astype(np.float32), V


def write_mm(path, doc_ptr, word_id, count, V):
    """CSR corpus -> the same file format (`MmCorpus.serialize`), so that the reference's scripts can read it."""
    doc_ptr = np.asarray(doc_ptr, np.int64)
    D = len(doc_ptr) - 1
    rows = np.repeat(np.arange(1, D + 1), np.diff(doc_ptr))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{D} {int(V)} {len(word_id)}\n")
        for r, w, c in zip(rows.tolist(), np.asarray(word_id).tolist(), np.asarray(count).tolist()):
            f.write(f"{r} {w + 1} {int(c) if float(c).is_integer() else repr(float(c))}\n")
