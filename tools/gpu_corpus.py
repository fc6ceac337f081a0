"""Timing of the device corpus sampler at BASELINE config 3's shape (development tool, GPU box):
    python tools/gpu_corpus.py [--docs 100000] [--V 10000] [--K 50] [--n-words 150]"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=100000)
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--V", type=int, default=10000)
    ap.add_argument("--n-words", type=int, default=150)
    a = ap.parse_args()
    import torch
    from strutopy_b200 import _lib
    rng = np.random.default_rng(1)
    theta = rng.dirichlet(np.ones(a.K), a.docs)
    beta = rng.dirichlet(np.full(a.V, 0.05), a.K)
    dev = torch.device("cuda", 0)
    ctx = _lib.Context(a.K, a.V, 1, 0)
    th, be = torch.from_numpy(theta).to(dev), torch.from_numpy(beta).to(dev)
    ptr = torch.empty(a.docs + 1, dtype=torch.int64, device=dev)
    ids = torch.empty(a.docs * a.n_words, dtype=torch.int32, device=dev)
    cnt = torch.empty(a.docs * a.n_words, dtype=torch.float32, device=dev)
    nnz = C.c_int64()
    st = torch.cuda.current_stream(dev).cuda_stream
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        _lib.check(ctx.handle, _lib.load().stm_sample_corpus(ctx.handle, a.docs, a.n_words, th.data_ptr(), be.data_ptr(),
                                                             C.c_uint64(12345), ptr.data_ptr(), ids.data_ptr(),
                                                             cnt.data_ptr(), C.byref(nnz), st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"rep {rep}: {ms:.2f} ms for {a.docs} documents x {a.n_words} tokens "
              f"({a.docs * a.n_words / ms / 1e6:.2f} G tokens/s), nnz {nnz.value}, mean distinct words {nnz.value / a.docs:.1f}")
    t = time.perf_counter()
    p = theta[:2000] @ beta
    for d in range(2000):
        rng.multinomial(a.n_words, p[d])
    t = time.perf_counter() - t
    print(f"host NumPy (the reference's dense theta @ beta + rng.multinomial per document): {2000 / t:.0f} docs/s on 1 core")


if __name__ == "__main__":
    main()
