"""Parity sweep (development tool, GPU box): the CUDA E-step through stm_estep_host against the C oracle over
several seeds / shapes / states; prints ELBO relative error, the share of documents whose eta differs by more
than 1e-6, and BFGS status / nit / repair mismatches.  python tools/gpu_parity_sweep.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import random_init_beta, synthetic_corpus  # noqa: E402
from oracle import c_oracle  # noqa: E402
from strutopy_b200 import _lib  # noqa: E402


def main():
    for K, V, D, nw in ((5, 500, 4000, 100), (20, 5000, 4000, 150), (50, 10000, 6000, 150), (100, 20000, 2000, 150)):
        for seed in (1, 2, 3):
            ptr, ids, cnt, X, _ = synthetic_corpus(D, V, K, n_words=nw, seed=100 * K + seed)
            rng = np.random.default_rng(seed)
            beta = random_init_beta(K, V) if seed == 1 else rng.dirichlet(np.full(V, 0.05), K)
            beta = np.maximum(beta, 1e-30).astype(np.float32).astype(np.float64)
            sigma = np.eye(K - 1) * (20.0 if seed == 1 else 1.5) + (0.0 if seed == 1 else 0.2)
            siginv, ent = c_oracle.prologue(sigma)
            mu = np.zeros((D, K - 1)) if seed == 1 else rng.normal(0, 0.5, size=(D, K - 1))
            eta0 = np.zeros((D, K - 1)) if seed == 1 else rng.normal(0, 0.5, size=(D, K - 1))
            ref = c_oracle.estep(ptr, ids, cnt, beta, mu, siginv, ent, eta0, nthreads=os.cpu_count() or 4)
            ctx = _lib.Context(K, V, 1)
            ctx.set_corpus(ptr, ids, cnt)
            o = ctx.estep_host(beta, mu, siginv, ent, eta0)
            ctx.close()
            d = np.abs(o["eta"] - ref["eta"]).max(axis=1)
            print("K=%d V=%d D=%d state=%s: ELBO rel %.1e | share of docs with |d eta| > 1e-6: %.5f (max %.1e) | "
                  "status/nit/repair mismatches %d/%d/%d (repair rate %.2f) | sigma_ss rel %.1e" % (
                      K, V, D, "init" if seed == 1 else "random%d" % seed,
                      abs(o["bound"] - ref["bound"]) / abs(ref["bound"]), float((d > 1e-6).mean()), float(d.max()),
                      int((o["status"] != ref["status"]).sum()), int((o["nit"] != ref["nit"]).sum()),
                      int((o["repair"] != ref["repair"]).sum()), float(np.mean(ref["repair"] > 0)),
                      float(np.abs(o["sigma_ss"] - ref["sigma_ss"]).max() / np.abs(ref["sigma_ss"]).max())), flush=True)


if __name__ == "__main__":
    main()
