"""Development tool (GPU box): ELBO traces of the K=50 spectral fixture — CUDA path vs the live reference's trace
(200-document cut) and vs the C oracle's EM at D=2000 (beta rounded to fp32 after every M-step, and not)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, unpack_corpus  # noqa: E402
from oracle import c_oracle, stm_numpy  # noqa: E402

np.set_printoptions(linewidth=220, precision=2)


def rel(a, b):
    n = min(len(a), len(b))
    return np.abs((np.asarray(a[:n]) - np.asarray(b[:n])) / np.asarray(b[:n]))


def main():
    from strutopy_b200 import STM
    g = load_golden("em_k50.npz")
    K, V, cut = int(g["K"]), int(g["V"]), int(g["cut"])
    beta0 = g["beta0"].astype(np.float64)
    nt = os.cpu_count() or 4
    run = lambda *a, **k: c_oracle.estep(*a, nthreads=nt, **k)  # noqa: E731
    for tag, n_docs in (("cut", cut), ("full", None)):
        ptr, ids, cnt = unpack_corpus(g, n_docs)
        X = g["X"][:len(ptr) - 1]
        m = STM((ptr, ids, cnt), range(V), False, K, X, False, 25, 0, 1e-5, init_type="random", model_type="STM")
        m.beta = beta0
        m.expectation_maximization(saving=False)
        r32 = stm_numpy.em(ptr, ids, cnt, beta0, X, n_iter=25, estep_fn=run, round_beta32=True)
        r64 = stm_numpy.em(ptr, ids, cnt, beta0, X, n_iter=25, estep_fn=run)
        print(f"[{tag}] D={len(ptr) - 1}")
        if tag == "cut":
            print("  gpu      vs live reference:", rel(m.last_bounds, g["cut_bounds"]))
            print("  oracle64 vs live reference:", rel(r64["bounds"], g["cut_bounds"]))
            print("  oracle32 vs live reference:", rel(r32["bounds"], g["cut_bounds"]))
        print("  gpu      vs oracle32      :", rel(m.last_bounds, r32["bounds"]))
        print("  gpu      vs oracle64      :", rel(m.last_bounds, r64["bounds"]))
        print("  oracle32 vs oracle64      :", rel(r32["bounds"], r64["bounds"]))


if __name__ == "__main__":
    main()
