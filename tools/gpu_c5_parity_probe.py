"""Development tool (GPU box): BASELINE config 5's shape (K=100, V=20k, 2 content aspects, kappa update) after two EM
iterations — the CUDA E-step vs the C oracle vs the NumPy port on a document sample; dumps the documents whose BFGS
iteration count differs so that they can be studied offline."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import c_oracle, stm_numpy  # noqa: E402


def main():
    import torch
    from strutopy_b200 import STM, _lib
    K, V, D, A = 100, 20000, int(os.environ.get("DOCS", "20000")), 2
    ptr, ids, cnt, X = bench.make_corpus(D, V, K)
    aspect = (np.arange(D) % A).astype(np.int32)
    m = STM((ptr, ids, cnt), range(V), True, K, X, True, 10 ** 9, 0, 0.0, init_type="spectral", model_type="STM",
            A=A, beta_index=aspect, lda_beta=False)
    for _ in range(2):
        m._estep_device(); m._reduce_and_bound(); m._mstep_device()
    snap = dict(beta=np.array(m.beta), mu=np.array(m.mu), sigma=np.array(m.sigma), eta0=np.array(m.eta))
    siginv, ent = bench.host_prologue(snap["sigma"])
    beta_r = snap["beta"].astype(np.float32).astype(np.float64)
    ctx = _lib.Context(K, V, A)
    ctx.set_corpus(ptr, ids, cnt, aspect)
    g = ctx.estep_host(beta_r, snap["mu"], siginv, ent, snap["eta0"])
    o = c_oracle.estep(ptr, ids, cnt, beta_r, snap["mu"], siginv, ent, snap["eta0"], aspect=aspect, nthreads=os.cpu_count())
    sel = np.sort(np.random.default_rng(5).choice(D, size=512, replace=False))
    pool = bench.PortPool(os.cpu_count())
    _, port = pool.run(pool.jobs(ptr, ids, cnt, beta_r, snap["mu"], siginv, ent, snap["eta0"], sel, aspect))
    pool.close()
    for name, a, b in (("gpu vs oracle (all)", g, o),):
        d = np.abs(a["eta"] - b["eta"]).max(axis=1)
        print(f"{name}: nit mismatches {(a['nit'] != b['nit']).sum()} / {D}, status {(a['status'] != b['status']).sum()}, "
              f"eta max {d.max():.2e}, docs > 1e-6: {(d > 1e-6).sum()}, ELBO rel {abs(a['bound'] - b['bound']) / abs(b['bound']):.2e}")
    for name, a in (("gpu", g), ("oracle", o)):
        d = np.abs(a["eta"][sel] - port["eta"]).max(axis=1)
        print(f"{name} vs port (sample 512): nit mismatches {(a['nit'][sel] != port['nit']).sum()}, eta max {d.max():.2e}, "
              f"docs > 1e-6: {(d > 1e-6).sum()}, doc_bound max rel {np.max(np.abs(a['doc_bound'][sel] - port['doc_bound']) / np.abs(port['doc_bound'])):.2e}")
    bad = np.flatnonzero(g["nit"] != o["nit"])[:24]
    print("nit gpu / oracle of the first mismatching documents:", list(zip(g["nit"][bad], o["nit"][bad])))
    print("beta: min positive", beta_r[beta_r > 0].min(), "zeros", int((beta_r == 0).sum()), "of", beta_r.size,
          "| beta < 1e-30:", int((beta_r < 1e-30).sum()))
    out = {}
    for j, d in enumerate(bad):
        lo, hi = ptr[d], ptr[d + 1]
        out[f"d{j}_ids"], out[f"d{j}_cnt"] = ids[lo:hi], cnt[lo:hi]
        out[f"d{j}_beta"] = beta_r[aspect[d]][:, ids[lo:hi]]
        out[f"d{j}_mu"], out[f"d{j}_eta0"] = snap["mu"][d], snap["eta0"][d]
        out[f"d{j}_eta_gpu"], out[f"d{j}_eta_oracle"] = g["eta"][d], o["eta"][d]
        out[f"d{j}_nit"] = np.array([g["nit"][d], o["nit"][d]])
    out["siginv"], out["ent"] = siginv, np.float64(ent)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "c5_mismatch_docs.npz"), **out)


if __name__ == "__main__":
    main()
