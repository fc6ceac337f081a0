"""Timing of the spectral initialisation at BASELINE config 3's shape (development tool, GPU box):
    python tools/gpu_spectral.py [--docs 100000] [--V 10000] [--K 50] [--maxV 5000]
Prints CUDA-event times of the two phases and sanity figures of the result."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=100000)
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--V", type=int, default=10000)
    ap.add_argument("--maxV", type=int, default=5000)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from strutopy_b200 import _lib
    from strutopy_b200.spectral import keep_list
    ptr, ids, cnt, X = bench.make_corpus(a.docs, a.V, a.K)
    totals = np.bincount(ids, weights=cnt.astype(np.float64), minlength=int(ids.max()) + 1)
    keep, wkeep = keep_list(totals, a.maxV)
    n = len(keep)
    keep32 = np.ascontiguousarray(keep, np.int32)
    ctx = _lib.Context(a.K, a.V, 1, 0)
    ctx.set_corpus(ptr, ids, cnt)
    L, h = _lib.load(), ctx.handle
    dev = torch.device("cuda", 0)
    gram = torch.empty(n * n + n, dtype=torch.float64, device=dev)
    beta = torch.empty((a.K, a.V), dtype=torch.float64, device=dev)
    anchors = np.zeros(a.K, np.int32)
    st = torch.cuda.current_stream(dev).cuda_stream
    for rep in range(a.reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        ev[0].record()
        _lib.check(h, L.stm_spectral_gram(h, n, _lib.hp(keep32), gram.data_ptr(), st))
        ev[1].record()
        _lib.check(h, L.stm_spectral_finish(h, n, _lib.hp(keep32), _lib.hp(wkeep), gram.data_ptr(), beta.data_ptr(),
                                            _lib.hp(anchors), st))
        ev[2].record()
        torch.cuda.synchronize()
        print(f"rep {rep}: gram {ev[0].elapsed_time(ev[1]):.1f} ms (D={a.docs}, kept {n} of V={a.V}: "
              f"{a.docs * n * n / ev[0].elapsed_time(ev[1]) / 1e9:.1f} TFLOP/s fp64 syrk-equivalent), "
              f"anchors + recover {ev[1].elapsed_time(ev[2]):.1f} ms")
    b = beta.cpu().numpy()
    print("anchors (first 10)", keep[anchors][:10], "row sums", b.sum(1)[:3], "min", b.min(), "nonzero frac", (b > 0.001 / a.V * 1.0001 / b.sum()).mean())


if __name__ == "__main__":
    main()
