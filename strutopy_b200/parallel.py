"""Data-parallel plumbing: documents are sharded across ranks (one process per GPU); the only
exchange per EM iteration is ONE sum-all-reduce of the packed sufficient-statistics buffer
(include/stm_b200.h "packed sufficient-statistics buffer").  The reference has no
distributed path (its documents loop is serial, stm.py:519) — this module is new.

Host-only logic (NumPy + torch.distributed); runs on CPU with the gloo backend in the tests.
"""
import numpy as np


def shard_bounds(doc_ptr, world):
    """Contiguous document ranges [(lo, hi)] * world, balanced by nnz (per-document cost is
    proportional to the number of distinct words), never splitting a document."""
    doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
    D = doc_ptr.shape[0] - 1
    if world <= 1:
        return [(0, D)]
    nnz = int(doc_ptr[-1])
    # weight = nnz + a per-document constant so that empty / tiny documents still spread out
    w = (doc_ptr[1:] - doc_ptr[:-1]).astype(np.float64) + (nnz / max(D, 1)) * 0.25 + 1e-9
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        c = int(np.searchsorted(cum, target, side="left"))
        c = min(max(c, cuts[-1]), D)
        cuts.append(c)
    cuts.append(D)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def stats_layout(A, V, TS, K, p):
    """Python mirror of stm_stats_layout (include/stm_b200.h): segment offsets + total length."""
    K1 = K - 1
    sizes = [A * V * TS, K1 * K1, 1, 1, K1, p, p * p, p * K1, K1 * K1]
    off = [0]
    for s in sizes:
        off.append(off[-1] + s)
    return off


def pack_stats(off, beta_ss_t, sigma_ss, bound, n_docs, eta, X):
    """Build the packed fp64 statistics vector of one shard on the host (test helper / spec):
    segments as documented in include/stm_b200.h."""
    out = np.zeros(off[9])
    out[off[0]:off[1]] = np.asarray(beta_ss_t, dtype=np.float64).reshape(-1)
    out[off[1]:off[2]] = np.asarray(sigma_ss, dtype=np.float64).reshape(-1)
    out[off[2]] = bound
    out[off[3]] = n_docs
    eta = np.asarray(eta, dtype=np.float64)
    X = np.asarray(X, dtype=np.float64).reshape(eta.shape[0], -1)
    out[off[4]:off[5]] = eta.sum(axis=0)
    out[off[5]:off[6]] = X.sum(axis=0)
    out[off[6]:off[7]] = (X.T @ X).reshape(-1)
    out[off[7]:off[8]] = (X.T @ eta).reshape(-1)
    out[off[8]:off[9]] = (eta.T @ eta).reshape(-1)
    return out


def allreduce_stats(stats, dist=None):
    """The one collective of an EM iteration.  `stats` is a torch tensor (CUDA with NCCL, CPU with
    gloo); reduced in place and returned."""
    if dist is None:
        import torch.distributed as dist  # noqa: PLW0642
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def mstep_from_stats(off, stats, X_local, K, p, model="STM", sigprior=0.0):
    """Host fp64 statement of what stm_mstep computes from the REDUCED statistics (spec + CPU test
    oracle for the multi-rank algebra): centred min-norm OLS with the intercept dropped
    (stm.py:691-706 via sklearn LinearRegression), Sigma from expanded moments (stm.py:723-728)."""
    K1 = K - 1
    N = stats[off[3]]
    sigma_ss = stats[off[1]:off[2]].reshape(K1, K1)
    sum_eta = stats[off[4]:off[5]]
    sum_x = stats[off[5]:off[6]]
    xtx = stats[off[6]:off[7]].reshape(p, p)
    xte = stats[off[7]:off[8]].reshape(p, K1)
    ete = stats[off[8]:off[9]].reshape(K1, K1)
    if model == "CTM":
        mean = sum_eta / N
        mu = np.repeat(mean[None, :], X_local.shape[0], axis=0)
        cov = ete - np.outer(mean, sum_eta)
        gamma = None
    else:
        G = xtx - np.outer(sum_x / N, sum_x)
        R = xte - np.outer(sum_x / N, sum_eta)
        lam, vec = np.linalg.eigh(G)
        keep = (lam > 1e-12 * lam.max()) & (lam > 0)
        inv = np.where(keep, 1.0 / np.where(keep, lam, 1.0), 0.0)
        gamma_t = vec @ (inv[:, None] * (vec.T @ R))  # p x K1
        mu = X_local @ gamma_t
        T = gamma_t.T @ xte
        cov = ete - T.T - T + gamma_t.T @ xtx @ gamma_t
        gamma = gamma_t.T
    sigma = (cov + sigma_ss) / N
    sigma = np.diag(np.diag(sigma)) * sigprior + (1 - sigprior) * sigma
    return dict(mu=mu, gamma=gamma, sigma=sigma)
