"""CPU check of kernel A's curvature certificate on many states (test infrastructure: uses oracle/).

The C oracle replays every line search in full, as SciPy does, and - with stm_oracle_shortcut_check enabled - evaluates
the kernel's rule on the way: it counts the searches in which the certificate held and those of them that went on to
accept a step (must be 0), and the trials the kernel does not make.  Output: profiles/r02d_shortcut_check.txt.
usage: python tools/oracle_shortcut_check.py"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import c_oracle, stm_numpy  # noqa: E402

NT = os.cpu_count() or 4
TOTAL = dict(docs=0, states=0, viol=0)


def check(tag, ptr, ids, cnt, beta, mu, sigma, eta, aspect=None):
    siginv, ent = stm_numpy.prologue(sigma)
    c_oracle.shortcut_check(True)
    try:
        r = c_oracle.estep(ptr, ids, cnt, beta, mu, siginv, ent, eta, aspect=aspect, nthreads=NT)
    finally:
        c = c_oracle.shortcut_check(False)
    D = len(ptr) - 1
    viol = c["cert_w1_accept_after"] + c["cert_zoom_accept_after"]
    TOTAL["docs"] += D; TOTAL["states"] += 1; TOTAL["viol"] += viol
    print(f"{tag:44s} D={D:5d} status0/2 {np.mean(r['status'] == 0):.2f}/{np.mean(r['status'] == 2):.2f} nit {r['nit'].mean():5.2f} | "
          f"certificate held: DCSRCH {c['cert_w1_fired']:5d} _zoom {c['cert_zoom_fired']:5d}  trials not made/doc "
          f"{c['cert_trials_skipped'] / D:5.1f}  acceptances after it {viol}", flush=True)
    assert viol == 0, tag


def synthetic(D, V, K, init, warmups, n_words=150, seed=12345, A=1):
    ptr, ids, cnt, X = bench.make_corpus(D, V, K, n_words=n_words, seed=seed)
    asp = (np.arange(D) % A).astype(np.int32) if A > 1 else None
    for wu in warmups:
        snap = bench.cpu_snapshot_state(types.SimpleNamespace(init=init, warmup=wu), ptr, ids, cnt, X, asp, K, V, A, NT)
        check(f"K={K} V={V} {n_words} tokens A={A} {init} it{wu}", ptr, ids, cnt.astype(np.float64), snap["beta"],
              snap["mu"], snap["sigma"], snap["eta0"], aspect=asp)


def main():
    synthetic(1500, 10000, 50, "spectral", (0, 1, 3))
    synthetic(1500, 5000, 20, "random", (0, 2))
    synthetic(1000, 500, 5, "random", (0, 3))
    synthetic(600, 20000, 100, "spectral", (1,))
    synthetic(1200, 3000, 30, "random", (2,), seed=7, A=2)
    synthetic(500, 20000, 100, "random", (2,), seed=7, A=2)
    synthetic(2000, 300, 3, "random", (0, 2), n_words=20, seed=11)
    synthetic(2000, 1000, 8, "random", (0, 2), n_words=400, seed=11)
    synthetic(600, 800, 12, "random", (0, 1, 4), n_words=2000, seed=21)
    synthetic(1500, 200, 2, "random", (0, 1, 4), n_words=50, seed=21)
    synthetic(800, 4000, 40, "random", (0, 1, 4), n_words=8, seed=21)
    # the reference's shipped wiki corpus (tests/golden/wiki_corpus.npz), CTM, the reference's random init
    g = np.load(os.path.join(ROOT, "tests", "golden", "wiki_corpus.npz"))
    ptr, ids, cnt, V = g["doc_ptr"], g["word_id"], g["count"].astype(np.float64), int(g["V"])
    D = len(ptr) - 1
    for K in (10, 50, 70):
        rs = np.random.RandomState(123456)
        b = rs.gamma(0.1, 1, V * K).reshape(K, V)
        run = lambda *a, **k: c_oracle.estep(*a, nthreads=NT, **k)  # noqa: E731
        ref = stm_numpy.em(ptr, ids, cnt, b / b.sum(1, keepdims=True), np.zeros((D, 1)), n_iter=4, estep_fn=run,
                           round_beta32=True, keep_states=True, model="CTM")
        for t, st in enumerate(ref["states"]):
            check(f"wiki corpus K={K} it{t}", ptr, ids, cnt, st["beta"], st["mu"], st["sigma"], st["eta"])
    # adversarial: random eta / mu far from any optimum under tight and loose priors
    rng = np.random.default_rng(3)
    ptr, ids, cnt, _ = bench.make_corpus(1000, 3000, 25, seed=5)
    beta = bench.random_beta(25, 3000).astype(np.float32).astype(np.float64)
    for scale, sig in ((5.0, 0.01), (5.0, 100.0), (0.1, 1e-4), (20.0, 1.0)):
        eta = rng.normal(size=(1000, 24)) * scale
        mu = rng.normal(size=(1000, 24)) * scale
        check(f"adversarial eta, mu ~ {scale} N(0,1), Sigma = {sig} I", ptr, ids, cnt.astype(np.float64), beta, mu,
              np.eye(24) * sig, eta)
    print(f"TOTAL: {TOTAL['docs']} documents in {TOTAL['states']} states, {TOTAL['viol']} acceptances after a certificate")


if __name__ == "__main__":
    main()
