"""A/B timing of the E-step kernel (development tool).  Usage on the GPU box:
    python tools/gpu_perf.py [--iters 4] [--docs 100000] [--tune bfgs_slots=1,bfgs_warps=6]
Prints per-EM-iteration E-step kernel time (CUDA events), mean evaluations per document, ELBO."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--docs", type=int, default=100000)
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--V", type=int, default=10000)
    ap.add_argument("--tag", default="default")
    ap.add_argument("--init", default="random", choices=["random", "spectral"])
    ap.add_argument("--heldout", action="store_true", help="also time STM.eval_heldout on the training corpus")
    ap.add_argument("--tune", default="", help="stm_tune overrides, e.g. bfgs_slots=1,bfgs_warps=6")
    ap.add_argument("--slots-timing", action="store_true", help="the library is a -DSTM_SLOTS_TIMING=1 build")
    ap.add_argument("--lib", default="", help="path of an A/B build of the library (make variant)")
    a = ap.parse_args()
    import torch
    from strutopy_b200 import _lib as lib_
    if a.lib:
        lib_.LIB_PATH = os.path.abspath(a.lib)
        a.tag = a.lib
    if a.tune:
        lib_.DEFAULT_TUNE = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in a.tune.split(",")}
        a.tag = a.tune
    cache = f"/tmp/corpus_{a.docs}_{a.V}_{a.K}.npz"
    if os.path.exists(cache):
        z = np.load(cache)
        ptr, ids, cnt, X = z["ptr"], z["ids"], z["cnt"], z["X"]
    else:
        ptr, ids, cnt, X = bench.make_corpus(a.docs, a.V, a.K)
        np.savez(cache, ptr=ptr, ids=ids, cnt=cnt, X=X)
    from strutopy_b200 import STM
    m = STM((ptr, ids, cnt), range(a.V), False, a.K, X, False, 10 ** 9, 0, 0.0, init_type=a.init,
            model_type="STM", device=0)
    if a.init == "random":
        m.beta = bench.random_beta(a.K, a.V)
    out = []
    for it in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        m._estep_device()
        e1.record()
        b = m._reduce_and_bound()
        t = time.perf_counter()
        m._mstep_device()
        torch.cuda.synchronize()
        tm = time.perf_counter() - t
        d = m.doc_diagnostics()
        try:
            import ctypes as C
            from strutopy_b200 import _lib as L_
            fn = L_.load().stm_dbg_cycles
            arr = (C.c_ulonglong * 16)()
            fn(m._ctx.handle, arr)
            if a.slots_timing:
                rounds, fresh, memo = arr[8] or 1, arr[9] or 1, arr[10]
                names = ["refill", "evals", "scalar", "accept", "scalar2", "finalize"]
                tot = sum(arr[:6]) or 1
                out.append(f"      rounds {rounds}  fresh evals {fresh}  memo hits {memo}  cycles/round: " +
                           "  ".join(f"{nm} {c/rounds:.0f} ({c/tot*100:.0f}%)" for nm, c in zip(names, arr)) +
                           f"  | per fresh eval: exp+partials {arr[11]/fresh:.0f}  contraction {arr[12]/fresh:.0f}  log+reduce {arr[13]/fresh:.0f}")
            else:
                tot = sum(arr[:8]) or 1
                names = ["gather+a_k", "eval", "linesearch", "accept/Hupd", "theta", "colsum+hess", "chol", "bound+inv+nu",
                         "e:memo+max", "e:exp", "e:contract+lp", "e:log(prod)", "e:reduce", "e:lse+grad", "#steps", "#memo"]
                out.append("      cycles/doc/warp: " + "  ".join(f"{nm} {c/len(d['nfev']):.0f} ({c/tot*100:.0f}%)" for nm, c in zip(names, arr)))
        except AttributeError:
            pass
        nf = d['nfev']
        out.append(f"      nfev pct50/90/99/99.9/max {np.percentile(nf,50):.0f}/{np.percentile(nf,90):.0f}/{np.percentile(nf,99):.0f}/"
                   f"{np.percentile(nf,99.9):.0f}/{nf.max()}  nit max {d['nit'].max()}  sum nfev {nf.sum()}  status {np.bincount(d['status'])}")
        ka, kb = m._ctx.estep_kernel_ms()
        out.append(f"it{it}: estep {e0.elapsed_time(e1):7.2f} ms (A {ka:6.2f} + B {kb:5.2f})  nfev {d['nfev'].mean():5.1f}  nit {d['nit'].mean():4.2f} "
                   f"repair {np.mean(d['repair'] > 0):.2f}  mstep {tm*1e3:5.2f} ms  bound {b:.6f}")
    if a.heldout:
        m.eval_heldout((ptr, ids, cnt))   # warm-up (uploads the CSR again: timed below with the upload)
        import ctypes as C
        from strutopy_b200 import _lib as L_
        d_ptr = torch.from_numpy(np.ascontiguousarray(ptr, np.int64)).cuda()
        d_ids = torch.from_numpy(np.ascontiguousarray(ids, np.int32)).cuda()
        d_cnt = torch.from_numpy(np.ascontiguousarray(cnt, np.float32)).cuda()
        res = torch.empty(a.docs + 1, dtype=torch.float64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            L_.check(m._ctx.handle, L_.load().stm_heldout(m._ctx.handle, a.docs, d_ptr.data_ptr(), d_ids.data_ptr(),
                                                           d_cnt.data_ptr(), m._ptr("theta"), m._ptr("beta_t"),
                                                           res.data_ptr(), res.data_ptr() + 8 * a.docs, m._stream()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        nnz = int(ptr[-1])
        gb = (nnz * (8 + 4 * a.K) + a.docs * (8 * a.K + 16)) / 1e9
        out.append(f"heldout: {ms:.3f} ms per call over {a.docs} documents / {nnz} held-out words "
                   f"({a.docs / ms * 1e3 / 1e6:.1f} M docs/s, {gb / ms * 1e3:.0f} GB/s algorithmic), mean ll {float(res[-1]):.6f}")
    print(f"== {a.tag}")
    print("\n".join(out), flush=True)


if __name__ == "__main__":
    main()
