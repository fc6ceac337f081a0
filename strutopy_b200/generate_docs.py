"""Synthetic corpora on the GPU — the reference's `modules/generate_docs.py`.

`CorpusCreation` keeps the reference's constructor, attributes and methods
(/root/reference/src/modules/generate_docs.py:99-417).  The parameter draws (beta, gamma, metadata, eta,
theta) are the reference's own NumPy calls on the host (they are O(D K) and cheap); what moves to the
device is `sample_documents` (generate_docs.py:293-316): the reference materialises theta @ beta (D x V
dense, 8 GB at D=100k, V=10k) and draws one `rng.multinomial` per document in a Python loop; here
`stm_sample_corpus` (include/stm_b200.h) draws every token's (topic, word) pair in one kernel and leaves
the corpus as CSR.  Same distribution, different random stream (Philox4x32-10, reproducible from `seed`).

Differences from the reference, all deliberate:
  * `documents` is built lazily from the CSR triple `csr` (a list of 100k x 150 tuples costs more than
    the sampling); `STM(...)` accepts `corpus.csr` directly;
  * `dictionary` is a plain {id: str(id)} dict (gensim is not a dependency; STM only uses len() and
    item lookup, stm.py:375, 1193);
  * eta is drawn per document with the reference's call up to 50 000 documents and in one vectorised
    draw beyond (same distribution);
  * `display_props` (matplotlib) is not provided.
There is no CPU fallback for the sampling.
"""
import ctypes as C
import logging

import numpy as np

from . import _lib
from .heldout import cut_in_half as _cut_in_half, split_corpus as _split_docs

logger = logging.getLogger(__name__)


def stable_softmax(x):
    """generate_docs.py:20-24"""
    xshift = x - np.max(x)
    exps = np.exp(xshift)
    return exps / np.sum(exps)


def sample_corpus(theta, beta, n_words, seed=12345, device=0):
    """doc_d ~ Multinomial(n_words, theta_d beta) for every row of theta, on the GPU.
    -> (doc_ptr int64 [D+1], word_id int32 [nnz], count float32 [nnz]), ids ascending within a document."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("strutopy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    if theta.ndim != 2 or beta.ndim != 2 or theta.shape[1] != beta.shape[0]:
        raise ValueError("theta must be (D, K) and beta (K, V)")
    D, K = theta.shape
    V = beta.shape[1]
    dev = torch.device("cuda", int(device))
    ctx = _lib.Context(K, V, 1, int(device))
    try:
        th = torch.from_numpy(theta).to(dev)
        be = torch.from_numpy(beta).to(dev)
        ptr = torch.empty(D + 1, dtype=torch.int64, device=dev)
        ids = torch.empty(D * int(n_words), dtype=torch.int32, device=dev)
        cnt = torch.empty(D * int(n_words), dtype=torch.float32, device=dev)
        nnz = C.c_int64()
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(ctx.handle, _lib.load().stm_sample_corpus(
            ctx.handle, D, int(n_words), th.data_ptr(), be.data_ptr(), C.c_uint64(int(seed) & (2 ** 64 - 1)),
            ptr.data_ptr(), ids.data_ptr(), cnt.data_ptr(), C.byref(nnz), st))
        n = int(nnz.value)
        return ptr.cpu().numpy(), ids[:n].cpu().numpy(), cnt[:n].cpu().numpy()
    finally:
        ctx.close()


def renumber_by_first_appearance(word_id):
    """The reference renumbers words in order of first appearance while it samples
    (`self.new_ids`, generate_docs.py:299-315).  -> (new ids, old id of every new id)"""
    uniq, first = np.unique(word_id, return_index=True)
    order = np.argsort(first, kind="stable")
    new_of_old = np.empty(len(uniq), dtype=np.int64)
    new_of_old[order] = np.arange(len(uniq))
    return new_of_old[np.searchsorted(uniq, word_id)].astype(np.int32), uniq[order]


def _alpha_vector(alpha, K):
    """Dirichlet concentration of the LDA data-generating process (generate_docs.py:139-151)."""
    if isinstance(alpha, np.ndarray):
        return alpha
    ranks = np.arange(1, K + 1)
    named = {"symmetric": np.full(K, 1.0 / K), "asymmetric": 1.0 / (ranks + np.sqrt(ranks))}
    return named[alpha] if isinstance(alpha, str) and alpha in named else np.repeat(alpha, K)


def _softmax_rows(eta):
    """theta = softmax([eta, 0]) row by row, max-shifted like the reference's stable_softmax (generate_docs.py:20-24, 268-271)."""
    full = np.concatenate([eta, np.zeros((eta.shape[0], 1))], axis=1)
    e = np.exp(full - full.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


class CorpusCreation:
    """generate_docs.py:28-137.  Extra keywords: `device`, `seed` (of the device sampler).

    The constructor consumes the two random streams in the reference's order — `default_rng(12345)` for beta,
    the metadata and the LDA thetas, NumPy's legacy global stream for gamma and eta (generate_docs.py:129-137) — so a
    caller that seeds `np.random` as the reference's scripts do (04_create_synthetic_corpora.py:45-47) gets the
    reference's beta, gamma, metadata and (up to 50 000 documents) eta."""

    def __init__(self, n_topics, n_docs, n_words, V, level, treatment=False, alpha="symmetric", dgp="STM",
                 metadata=None, alpha_treatment=None, beta=None, theta=None, gamma=None, device=0, seed=12345):
        self.K, self.n_docs, self.n_words, self.V = n_topics, n_docs, n_words, V
        self.dgp, self.level, self.treatment = dgp, level, treatment
        self.device, self.seed = device, seed
        self.csr = None
        self._documents = None
        self.rng = np.random.default_rng(12345)
        self.init_alpha(alpha, alpha_treatment, theta)
        self.word_topic_dist(beta)
        self.init_gamma(gamma)
        self.set_metadata(metadata)
        self.init_eta()
        self.init_theta(theta)

    # ---- parameter draws (generate_docs.py:139-271): same distributions, same order of RNG calls -------
    def init_alpha(self, alpha, alpha_treatment, theta):
        self.alpha = _alpha_vector(alpha, self.K)
        if not np.any(self.alpha) and theta is None:
            raise AssertionError("Either alpha or theta needs to be specified for generating documents.")
        if self.treatment:
            self.init_treatment(alpha_treatment)

    def init_treatment(self, alpha_treatment):
        if alpha_treatment is None:
            raise AssertionError("If treatment == True, the effect needs to be specified by alpha_treatment")
        if isinstance(alpha_treatment, np.ndarray):
            self.alpha_treatment = alpha_treatment
        else:
            auto = {"auto-linear": lambda a: np.flip(a), "auto-nonlinear": lambda a: np.exp(a)}
            if alpha_treatment in auto:
                self.alpha_treatment = auto[alpha_treatment](self.alpha)

    def word_topic_dist(self, beta):
        # K draws from Dirichlet(0.05 * 1_V) unless the caller brings a topic-word matrix (generate_docs.py:172-184)
        self.beta = np.asarray(beta) if beta is not None else self.rng.dirichlet(np.full(self.V, 0.05), size=self.K)

    def init_gamma(self, gamma, mean=None):
        # prevalence coefficients (K-1) x level around a N(., 0.001 I) draw of their mean (generate_docs.py:186-203)
        if gamma is not None:
            self.gamma = gamma
            return
        small = 0.001 * np.eye(self.level)
        centre = np.random.standard_normal(self.level) if mean is None else mean
        centre = np.random.multivariate_normal(centre, small)
        self.gamma = np.random.multivariate_normal(centre, small, self.K - 1)

    def set_metadata(self, metadata, metadata_levels=[0, 1]):
        if metadata is not None:
            assert metadata.shape == (self.n_docs, self.level), "Unexpected metadata shape provided."
            self.metadata = metadata
        else:   # iid levels per document and covariate (generate_docs.py:205-219)
            self.metadata = self.rng.choice(metadata_levels, size=(int(self.n_docs), self.level), replace=True, p=None)

    def init_eta(self):
        # eta_d ~ N(x_d gamma', 0.001 I) (generate_docs.py:221-228); beyond 50k documents in one vectorised draw
        centre = self.metadata @ self.gamma.T
        if self.n_docs > 50000:
            self.eta = centre + np.sqrt(0.001) * np.random.standard_normal(centre.shape)
        else:
            cov = 0.001 * np.eye(self.K - 1)
            self.eta = np.stack([np.random.multivariate_normal(row, cov) for row in centre])

    def init_theta(self, theta):
        if self.dgp == "STM":
            self.map_eta(eta=self.eta)
        elif self.dgp == "LDA" and theta is None:
            half = int(self.n_docs / 2)
            if self.treatment:
                self.theta = self.rng.dirichlet(self.alpha, size=half)
                self.theta_treatment = self.rng.dirichlet(self.alpha_treatment, size=half)
            else:
                self.theta = self.rng.dirichlet(self.alpha, size=self.n_docs)
        else:
            self.theta = np.array(theta)
            assert self.theta.ndim == 2, "theta needs to be a 2D numpy array"

    def map_eta(self, eta):
        self.theta = _softmax_rows(np.asarray(eta))

    # ---- sampling: on the device (generate_docs.py:273-316) -----------------------------------------
    def generate_documents(self, remove_terms=True, dictionary=True, display_props=False):
        logger.info(f"Create corpus for K={self.K} topics.")
        self.sample_documents()
        if remove_terms:
            self.remove_infrequent_terms()
        if dictionary:
            self.create_dictionary()
        if display_props:
            raise NotImplementedError("display_props needs matplotlib; not part of the accelerated path")

    def _doc_theta(self):
        if self.dgp == "LDA" and self.treatment == True:  # noqa: E712  (lda_probs, generate_docs.py:318-328)
            return np.concatenate((self.theta, self.theta_treatment), axis=0)
        return self.theta

    def sample_documents(self):
        theta = self._doc_theta()
        ptr, ids, cnt = sample_corpus(theta, self.beta, self.n_words, seed=self.seed, device=self.device)
        # ids in order of first appearance, like the reference's `new_ids`
        ids, self.new_ids_inverse = renumber_by_first_appearance(ids)
        self.new_ids = {int(o): i for i, o in enumerate(self.new_ids_inverse)}
        self.csr = (ptr, ids, cnt)
        self._documents = None

    @property
    def documents(self):
        """list of [(word_id, count), ...] per document (the reference's format), built on first use"""
        if self._documents is None and self.csr is not None:
            ptr, ids, cnt = self.csr
            il, cl = ids.tolist(), cnt.astype(np.int64).tolist()
            self._documents = [list(zip(il[ptr[d]:ptr[d + 1]], cl[ptr[d]:ptr[d + 1]])) for d in range(len(ptr) - 1)]
        return self._documents

    @documents.setter
    def documents(self, docs):
        self._documents = docs

    def remove_infrequent_terms(self):
        """generate_docs.py:330-346: compact the ids to the words that occur (ascending old id) and set V.
        (The reference zips the ascending new ids with each document's counts in document order; ids are
        kept attached to their own counts here.)"""
        ptr, ids, cnt = self.csr
        present = np.unique(ids)
        logger.info(f"removes {self.V - len(present)} words due to no occurence")
        self.csr = (ptr, np.searchsorted(present, ids).astype(np.int32), cnt)
        self._documents = None
        self.V = len(present)

    def create_dictionary(self):
        ids = self.csr[1]
        self.dictionary = {i: str(i) for i in range(int(ids.max()) + 1 if ids.size else 0)}

    def split_corpus(self, validation_set=False, document_completion=True, proportion=0.8):
        """generate_docs.py:381-399 (same split as heldout.split_corpus, results stored as attributes)"""
        train, half1, half2, validate = _split_docs(self.documents, validation_set, document_completion, proportion)
        self.train_docs = train
        n_train = len(train)
        self.test_docs = self.documents[n_train:] if validate is None else self.documents[n_train:len(self.documents) - len(validate)]
        if validate is not None:
            self.validate_docs = validate
        if document_completion:
            self.test_1_docs, self.test_2_docs = half1, half2

    def cut_in_half(self, doc_set):
        """generate_docs.py:401-417"""
        return _cut_in_half(doc_set)
