"""Quick GPU-side parity probe (development tool): CUDA E-step through the host C-ABI vs golden
fixtures and the C oracle.  Usage: python tools/gpu_check.py"""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("stm_lib", os.path.join(ROOT, "strutopy_b200", "_lib.py"))
lib = importlib.util.module_from_spec(spec)
spec.loader.exec_module(lib)
from conftest import load_golden, random_init_beta, synthetic_corpus  # noqa: E402
from oracle import c_oracle  # noqa: E402


def report(tag, o, ref):
    de = np.abs(o["eta"] - ref["eta"]).max(axis=1)
    print(f"[{tag}] bound gpu {o['bound']:.10f} ref {ref['bound']:.10f} rel {abs(o['bound']-ref['bound'])/abs(ref['bound']):.2e}"
          f" | eta max {de.max():.2e} med {np.median(de):.2e} n>1e-6 {(de > 1e-6).sum()}/{len(de)}"
          f" | theta {np.abs(o['theta']-ref['theta']).max():.2e}"
          f" | bss {np.abs(o['beta_ss']-ref['beta_ss']).max():.2e} sss {np.abs(o['sigma_ss']-ref['sigma_ss']).max():.2e}"
          f" | status eq {(o['status']==ref['status']).mean():.3f} nit eq {(o['nit']==ref['nit']).mean():.3f}"
          + (f" repair eq {(o['repair']==ref['repair']).mean():.3f}" if 'repair' in ref else ""), flush=True)


def golden_case(name, pfx):
    g = load_golden(name)
    K, V = int(g["K"]), int(g["V"])
    A = int(g["A"]) if "A" in g else 1
    ctx = lib.Context(K, V, A)
    ctx.set_corpus(g["doc_ptr"], g["word_id"], g["count"], g.get("aspect"))
    o = ctx.estep_host(g[pfx + "beta"].astype(np.float64), g[pfx + "mu"], g[pfx + "siginv"],
                       float(g[pfx + "sigmaentropy"]), g[pfx + "eta0"])
    ref = {k: g[pfx + k] for k in ("eta", "theta", "beta_ss", "sigma_ss", "status", "nit")}
    ref["bound"] = float(g[pfx + "bound"])
    report(f"{name}:{pfx}", o, ref)
    ctx.close()


def oracle_case(D, V, K, seed=1):
    ptr, ids, cnt, X, _ = synthetic_corpus(D, V, K, seed=seed)
    beta = random_init_beta(K, V).astype(np.float32).astype(np.float64)
    siginv, ent = c_oracle.prologue(np.eye(K - 1) * 20.0)
    mu = np.zeros((D, K - 1))
    eta0 = np.zeros((D, K - 1))
    t = time.time()
    ref = c_oracle.estep(ptr, ids, cnt, beta, mu, siginv, ent, eta0, nthreads=os.cpu_count())
    tc = time.time() - t
    ctx = lib.Context(K, V, 1)
    ctx.set_corpus(ptr, ids, cnt)
    ctx.estep_host(beta, mu, siginv, ent, eta0)
    t = time.time()
    o = ctx.estep_host(beta, mu, siginv, ent, eta0)
    tg = time.time() - t
    report(f"oracle D={D} V={V} K={K}", o, ref)
    print(f"    C oracle {tc:.3f}s ({D/tc:.0f} docs/s, {os.cpu_count()} threads)  GPU host-call {tg*1e3:.2f} ms ({D/tg:.0f} docs/s)", flush=True)
    ctx.close()


if __name__ == "__main__":
    golden_case("kat_small.npz", "it0_")
    golden_case("estep_K5.npz", "it0_")
    golden_case("estep_K5.npz", "it2_")
    golden_case("estep_K20.npz", "it0_")
    golden_case("estep_K20.npz", "it1_")
    golden_case("estep_K50.npz", "it0_")
    golden_case("estep_K50.npz", "it1_")
    golden_case("estep_content.npz", "it0_")
    oracle_case(2000, 2000, 20)
    oracle_case(4000, 5000, 50)
    oracle_case(1000, 3000, 100)
