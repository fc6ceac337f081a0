"""CPU tests of the host-side logic: C-ABI exports, corpus packing, design matrix, sharding, the packed
statistics algebra, and the world_size=2 all-reduce path on gloo."""
import ctypes
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, synthetic_corpus
from oracle import c_oracle, stm_numpy
from strutopy_b200 import _lib
from strutopy_b200.corpus import pack_corpus, word_counts
from oracle.stats_numpy import mstep_from_stats, pack_stats
from oracle.stats_numpy import stats_layout
from strutopy_b200.parallel import shard_bounds
from strutopy_b200.stm import design_matrix


def test_library_exports_every_declared_symbol():
    """libstm_b200.so loads without a GPU and exports exactly what include/stm_b200.h declares."""
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "stm_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(stm_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/stm_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    lib.stm_beta_stride.restype = ctypes.c_int
    for K, ts in ((3, 4), (5, 12), (20, 20), (50, 52), (70, 76), (100, 100), (128, 132), (32, 36)):
        assert lib.stm_beta_stride(K) == ts


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.StmError):
        _lib.Context(5, 100, 1)
    from strutopy_b200 import STM
    with pytest.raises(RuntimeError):
        STM([[(0, 1)]], range(4), False, 3, np.zeros((1, 1)), False, 2, 0, 1e-5, init_type="random")


def test_pack_corpus_and_wcounts():
    docs = [[(0, 3), (2, 1), (3, 2), (5, 4)], [], [(1, 2), (2, 2), (4, 1)]]
    ptr, ids, cnt = pack_corpus(docs)
    np.testing.assert_array_equal(ptr, [0, 4, 4, 7])
    np.testing.assert_array_equal(ids, [0, 2, 3, 5, 1, 2, 4])
    np.testing.assert_array_equal(cnt, [3, 1, 2, 4, 2, 2, 1])
    np.testing.assert_array_equal(word_counts(ptr, ids, cnt, 6), [3, 2, 3, 2, 1, 4])
    p2, i2, c2 = pack_corpus((ptr, ids, cnt))
    np.testing.assert_array_equal(p2, ptr)


def test_design_matrix_binary_and_onehot():
    X = np.array([[0], [1], [1]])
    np.testing.assert_array_equal(design_matrix(X), X.astype(float))
    np.testing.assert_array_equal(design_matrix(np.array([0, 1, 1])), [[0], [1], [1]])
    Xc = np.array([[2, 0], [0, 1], [1, 1]])
    d = design_matrix(Xc)
    assert d.shape == (3, 5)
    np.testing.assert_array_equal(d.sum(axis=1), [2, 2, 2])
    np.testing.assert_array_equal(d, stm_numpy.design_matrix(Xc))


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(0)
    ptr = np.concatenate([[0], np.cumsum(rng.integers(0, 300, 5000))])
    for world in (1, 2, 3, 8):
        b = shard_bounds(ptr, world)
        assert b[0][0] == 0 and b[-1][1] == 5000
        for (l0, h0), (l1, h1) in zip(b[:-1], b[1:]):
            assert h0 == l1
        nnz = np.array([ptr[h] - ptr[l] for l, h in b], dtype=float)
        assert nnz.max() <= 1.05 * nnz.mean() + 300
    assert shard_bounds(np.array([0, 5]), 4)[-1][1] == 1


def _shard_stats(g, pfx, lo, hi, off, TS):
    ptr = g["doc_ptr"]
    sl = slice(ptr[lo], ptr[hi])
    o = c_oracle.estep(ptr[lo:hi + 1] - ptr[lo], g["word_id"][sl], g["count"][sl],
                       g[pfx + "beta"].astype(np.float64), g[pfx + "mu"][lo:hi], g[pfx + "siginv"],
                       float(g[pfx + "sigmaentropy"]), g[pfx + "eta0"][lo:hi])
    K, V = o["beta_ss"].shape
    bt = np.zeros((V, TS))
    bt[:, :K] = o["beta_ss"].T
    return pack_stats(off, bt, o["sigma_ss"], o["bound"], hi - lo, o["eta"], g["X"][lo:hi]), o


def test_sharded_statistics_equal_unsharded_mstep():
    """Summing per-shard packed statistics and running the moment-based M-step reproduces the
    reference M-step on the whole corpus (the algebra stm_mstep implements on the device)."""
    g = load_golden("estep_K20.npz")
    pfx = "it1_"
    K, V = int(g["K"]), int(g["V"])
    TS = 20
    p = 1
    off = stats_layout(1, V, TS, K, p)
    D = len(g["doc_ptr"]) - 1
    total = np.zeros(off[9])
    for lo, hi in shard_bounds(g["doc_ptr"], 3):
        s, _ = _shard_stats(g, pfx, lo, hi, off, TS)
        total += s
    assert total[off[3]] == D
    assert abs(total[off[2]] - g[pfx + "bound"]) <= 1e-12 * abs(g[pfx + "bound"])
    r = mstep_from_stats(off, total, g["X"].reshape(D, -1).astype(float), K, p)
    np.testing.assert_allclose(r["gamma"], g[pfx + "m_gamma"], atol=1e-9)
    np.testing.assert_allclose(r["mu"], g[pfx + "m_mu"], atol=1e-9)
    np.testing.assert_allclose(r["sigma"], g[pfx + "m_sigma"], atol=1e-9)
    bss = total[off[0]:off[1]].reshape(V, TS)[:, :K].T
    np.testing.assert_allclose(bss, g[pfx + "beta_ss"], atol=1e-9)


def test_mstep_from_stats_rank_deficient_matches_lstsq():
    """one-hot designs are rank deficient after centring: min-norm solution like scipy lstsq(cond=1e-6)"""
    rng = np.random.default_rng(1)
    D, K = 300, 6
    cats = rng.integers(0, 4, size=D)
    X = design_matrix(cats)
    eta = rng.normal(size=(D, K - 1)) + X @ rng.normal(size=(4, K - 1))
    off = stats_layout(1, 8, 8, K, 4)
    st = pack_stats(off, np.zeros((8, 8)), np.zeros((K - 1, K - 1)), 0.0, D, eta, X)
    r = mstep_from_stats(off, st, X, K, 4)
    mu_ref, gamma_ref = stm_numpy.update_mu(eta, cats)
    np.testing.assert_allclose(r["gamma"], gamma_ref, atol=1e-9)
    np.testing.assert_allclose(r["mu"], mu_ref, atol=1e-9)
    np.testing.assert_allclose(r["sigma"], stm_numpy.update_sigma(eta, mu_ref, np.zeros((K - 1, K - 1))), atol=1e-9)


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import numpy as np, torch, torch.distributed as dist
from conftest import load_golden
from oracle.stats_numpy import stats_layout
from strutopy_b200.parallel import allreduce_stats, shard_bounds
from oracle.stats_numpy import mstep_from_stats
from test_host_logic import _shard_stats
dist.init_process_group('gloo', rank=int(os.environ['RANK']), world_size=2)
g = load_golden('estep_K5.npz'); pfx = 'it2_'
K, V, TS, p = int(g['K']), int(g['V']), 12, 1
off = stats_layout(1, V, TS, K, p)
lo, hi = shard_bounds(g['doc_ptr'], 2)[dist.get_rank()]
s, o = _shard_stats(g, pfx, lo, hi, off, TS)
t = torch.from_numpy(s)
allreduce_stats(t, dist)
tot = t.numpy()
D = len(g['doc_ptr']) - 1
assert tot[off[3]] == D
assert abs(tot[off[2]] - g[pfx + 'bound']) <= 1e-12 * abs(g[pfx + 'bound'])
r = mstep_from_stats(off, tot, g['X'].reshape(D, -1).astype(float)[lo:hi], K, p)
np.testing.assert_allclose(r['gamma'], g[pfx + 'm_gamma'], atol=1e-9)
np.testing.assert_allclose(r['mu'], g[pfx + 'm_mu'][lo:hi], atol=1e-9)
np.testing.assert_allclose(r['sigma'], g[pfx + 'm_sigma'], atol=1e-9)
dist.barrier(); dist.destroy_process_group()
print('rank', os.environ['RANK'], 'ok')
"""


def test_two_rank_allreduce_gloo(tmp_path):
    """world_size=2 on CPU/gloo: shard -> per-rank statistics -> ONE all-reduce -> replicated M-step."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for pr in procs:
        out, _ = pr.communicate(timeout=240)
        assert pr.returncode == 0, out


def test_new_fronts_fail_loudly_without_gpu():
    """spectral_init / sample_corpus / eval_heldout have no CPU fallback either."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from strutopy_b200.generate_docs import sample_corpus
    from strutopy_b200.spectral import spectral_init
    docs = (np.array([0, 2, 4]), np.array([0, 1, 0, 1], np.int32), np.array([1.0, 2.0, 2.0, 1.0]))
    with pytest.raises(RuntimeError):
        spectral_init(docs, 2, 2)
    with pytest.raises(RuntimeError):
        sample_corpus(np.full((3, 2), 0.5), np.full((2, 5), 0.2), 10)


def test_spectral_gram_statistics_add_over_shards():
    """What stm_spectral_gram all-reduces: Htilde'Htilde and diag(Hhat) are sums over documents, so the shards'
    statistics add up to the global ones and Q = sum - diag(sum) is the unsharded Q (stm.py:134-149)."""
    from oracle import spectral_numpy as sn
    ptr, ids, cnt, _, _ = synthetic_corpus(300, 200, 4, n_words=50, seed=5)
    wprob = sn.word_prob(ptr, ids, cnt)
    keep = sn.keep_order(wprob, 60)
    M = sn.dense_dtm(ptr, ids, cnt, width=len(wprob))[:, keep]

    def parts(Ms):
        wc = Ms.sum(axis=1, keepdims=True)
        div = wc * (wc - 1)
        Ht = Ms / np.sqrt(div)
        return Ht.T @ Ht, np.sum(Ms / div, axis=0)

    g_all, h_all = parts(M)
    g1, h1 = parts(M[:130])
    g2, h2 = parts(M[130:])
    np.testing.assert_allclose(g1 + g2, g_all, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(h1 + h2, h_all, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose((g1 + g2) - np.diag(h1 + h2), sn.gram(M), rtol=1e-11, atol=1e-14)


def test_first_appearance_renumbering():
    """generate_docs.py:299-315: ids are renumbered in order of first appearance"""
    from strutopy_b200.generate_docs import renumber_by_first_appearance
    new, old_of_new = renumber_by_first_appearance(np.array([7, 3, 7, 9, 3, 0, 9], np.int32))
    np.testing.assert_array_equal(new, [0, 1, 0, 2, 1, 3, 2])
    np.testing.assert_array_equal(old_of_new, [7, 3, 9, 0])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path = the NumPy/SciPy port on host cores) runs without a GPU
    and prints one JSON line with the keys the driver reads."""
    import json
    # a small shape, so that the whole arm (one warm-up EM iteration of the C oracle, the process pool, the one-core
    # figure) runs in seconds; the host-side spectral init it uses at full size is pinned in tests/test_spectral.py
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-docs-per-core", "1", "--docs", "3000", "--K", "20", "--V", "3000", "--init", "random"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "estep_docs_per_sec_K50_V10k" and line["unit"] == "docs/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "docs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["config"]["init"] == "random"
    assert line["cpu_baseline"]["one_core_as_shipped"]["cores"] == 1
    assert "same state" in line["cpu_baseline"]["sample"]


def test_matrix_market_corpus_round_trip(tmp_path):
    """gensim MmCorpus files (02_create_corpus.py:42, 03_fit_reference_model.py:43-46): write -> read gives the CSR
    back; the reference's own file layout (1-based, "D V nnz" size line, real counts) parses."""
    from strutopy_b200.corpus import read_mm, write_mm
    ptr, ids, cnt, _, _ = synthetic_corpus(40, 90, 3, n_words=30, seed=1)
    path = tmp_path / "BoW_corpus.mm"
    write_mm(path, ptr, ids, cnt, 90)
    p2, i2, c2, V = read_mm(path)
    assert V == 90
    np.testing.assert_array_equal(p2, ptr)
    np.testing.assert_array_equal(i2, ids)
    np.testing.assert_array_equal(c2, cnt)
    (tmp_path / "ref.mm").write_text("%%MatrixMarket matrix coordinate real general\n2 5 4                \n"
                                     "1 1 1\n1 3 2\n2 2 1\n2 5 3\n")
    p3, i3, c3, V3 = read_mm(tmp_path / "ref.mm")
    assert V3 == 5 and p3.tolist() == [0, 2, 4] and i3.tolist() == [0, 2, 1, 4] and c3.tolist() == [1.0, 2.0, 1.0, 3.0]
    w = load_golden("wiki_corpus.npz")   # the reference's shipped corpus survives the round trip too
    write_mm(tmp_path / "wiki.mm", w["doc_ptr"], w["word_id"], w["count"], int(w["V"]))
    p4, i4, c4, V4 = read_mm(tmp_path / "wiki.mm")
    np.testing.assert_array_equal(p4, w["doc_ptr"])
    np.testing.assert_array_equal(i4, w["word_id"])
    np.testing.assert_array_equal(c4, w["count"].astype(np.float32))


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under strutopy_b200/ may import, link or execute it, and the product
    must not carry a CPU fallback (only tests/, __graft_entry__.smoke() and bench.py's CPU legs use the oracle)."""
    pkg = os.path.join(ROOT, "strutopy_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                continue
            text = open(os.path.join(dirpath, fn)).read()
            for line in text.splitlines():
                code = line.split("#")[0].split("//")[0]
                assert not re.search(r"\b(import|from)\s+oracle\b", code), f"{fn}: {line.strip()}"
                assert "libstm_oracle" not in code and "oracle/_ref" not in code, f"{fn}: {line.strip()}"
    # importing the package pulls in nothing of the oracle
    out = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import strutopy_b200, "
                          "strutopy_b200.spectral, strutopy_b200.generate_docs, strutopy_b200.heldout; "
                          "print(sorted(m for m in sys.modules if m.split('.')[0] == 'oracle'))" % ROOT],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "[]", out.stdout + out.stderr
