#!/usr/bin/env python
"""bench.py — E-step docs/sec at K=50, V=10k (BASELINE.json metric) on synthetic corpora.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's NumPy/SciPy path on host cores

Workload (config.workload): BASELINE.json configs[2]/[3] — D=100k documents per GPU, V=10k, K=50, 150 tokens/doc from
the reference's DGP (generate_docs.py:180-316), spectral init (--init random: the reference's random init), 1
prevalence covariate.  A "step" is ONE EM iteration (E-step kernel pair + moments + the one all-reduce + M-step) over
the resident corpus; `value` = documents processed by all ranks / max-over-ranks device time.  `e2e` = the same
documents through the reference-facing host call stm_estep_host (fp64 host buffers in the reference's layouts, H2D +
D2H inside the timed region).  With N > 1 the line also carries a `strong` block: BASELINE config 4 (the SAME 100k
documents split over the N GPUs) measured in the same run, with a per-phase breakdown.  `--config c5` switches to
BASELINE config 5's shape (V=20k, K=100, 2 content aspects, kappa update on, D=500k/8 documents per GPU).

Both arms time the SAME state: the one the first timed E-step starts from (spectral init + W warm-up EM iterations).
The reference arm prepares it on the host (oracle/spectral_numpy.py + the C oracle's EM), then times the NumPy/SciPy
port (bit-identical to the reference on the golden fixtures) in a process pool created before the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "estep_docs_per_sec_K50_V10k"
UNIT = "docs/s"


def make_corpus(D, V, K, n_words=150, seed=12345, p=1):
    """Vectorised sampler with the reference DGP's distributions (generate_docs.py:180-316):
    beta_k ~ Dir(0.05), X ~ U{0,1}, eta ~ N(X gamma', 0.001 I), theta = softmax([eta, 0]),
    doc ~ Multinomial(n_words, theta beta) drawn as topic-then-word.  -> CSR + X."""
    rng = np.random.default_rng(seed)
    beta = rng.dirichlet(np.full(V, 0.05), K)
    cum = np.cumsum(beta, axis=1)
    cum[:, -1] = 1.0
    rng = np.random.default_rng(seed + 1000003 * (1 + int(os.environ.get("RANK", "0"))))
    X = rng.integers(0, 2, size=(D, p)).astype(np.float64)
    gamma = np.random.default_rng(seed + 7).normal(0.0, 1.0, size=(K - 1, p))
    eta = X @ gamma.T + rng.normal(0, np.sqrt(0.001), size=(D, K - 1))
    full = np.concatenate([eta, np.zeros((D, 1))], axis=1)
    theta = np.exp(full - full.max(1, keepdims=True))
    theta /= theta.sum(1, keepdims=True)
    tc = np.cumsum(theta, axis=1)
    tc[:, -1] = 1.0
    u = rng.random((D, n_words))
    z = np.empty((D, n_words), dtype=np.int32)
    step = max(1, (1 << 24) // (n_words * K))
    for lo in range(0, D, step):
        hi = min(D, lo + step)
        z[lo:hi] = (u[lo:hi, :, None] > tc[lo:hi, None, :]).sum(axis=2)
    np.minimum(z, K - 1, out=z)
    u2 = rng.random((D, n_words))
    w = np.empty((D, n_words), dtype=np.int64)
    zf, wf, uf = z.reshape(-1), w.reshape(-1), u2.reshape(-1)
    for k in range(K):
        sel = np.nonzero(zf == k)[0]
        wf[sel] = np.searchsorted(cum[k], uf[sel], side="left")
    np.minimum(wf, V - 1, out=wf)
    key = (np.repeat(np.arange(D, dtype=np.int64), n_words) * V + wf)
    uk, cnt = np.unique(key, return_counts=True)
    doc = uk // V
    ids = (uk % V).astype(np.int32)
    ptr = np.zeros(D + 1, dtype=np.int64)
    np.cumsum(np.bincount(doc, minlength=D), out=ptr[1:])
    return ptr, ids, cnt.astype(np.float32), X


def random_beta(K, V, seed=123456):
    """the reference's random init (stm.py:425-429) with its legacy-RNG seed"""
    rs = np.random.RandomState(seed)
    b = rs.gamma(0.1, 1, V * K).reshape(K, V)
    return b / b.sum(axis=1, keepdims=True)


def workload(args, world):
    """(K, V, documents per GPU, content aspects) of the chosen BASELINE config"""
    if args.config == "c5":
        return 100, 20000, (args.docs if args.docs else 500000 // 8), 2
    return args.K, args.V, (args.docs if args.docs else 100000), 1


def config_dict(args, world):
    """`config` of the JSON line — identical in both arms (the reference arm times the same workload and state)"""
    K, V, D, A = workload(args, world)
    name = "C5" if args.config == "c5" else "C3"
    extra = ", 2 content aspects (beta_index = d mod 2), kappa update on" if A > 1 else ""
    return {
        "workload": f"{name}: D={D}/GPU V={V} K={K}, 150 tokens/doc, reference DGP, {args.init} init + {args.warmup} "
                    f"warm-up EM iterations, 1 prevalence covariate{extra}; step = one EM iteration",
        "init": args.init, "docs_per_gpu": D, "V": V, "K": K, "A": A, "beta_storage": "fp32", "arithmetic": "fp64",
        "l2": "per-step working set (eta, mu, theta, corpus, beta_ss) > 250 MB exceeds the 126 MB L2",
        "parallelism": f"dp{world}: documents sharded, one NCCL all-reduce of the packed statistics per step",
    }


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.reasons, self.max_mhz = device, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------------
# CPU legs: the reference's NumPy/SciPy E-step (oracle/stm_numpy.py, bit-identical to
# /root/reference/src/modules/stm.py on the golden fixtures) in a process pool that outlives the timed region
# ----------------------------------------------------------------------------------------------------
def _ref_worker(args):
    from oracle import stm_numpy
    ptr, ids, cnt, beta, mu, siginv, ent, eta0, aspect = args
    t = time.perf_counter()
    try:
        # one BLAS / LAPACK thread per worker process: the pool already uses every core (at K=100 a multi-threaded
        # LAPACK in each of the workers oversubscribes the host 16x)
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            o = stm_numpy.estep(ptr, ids, cnt, beta, mu, siginv, ent, eta0, aspect=aspect)
    except ImportError:
        o = stm_numpy.estep(ptr, ids, cnt, beta, mu, siginv, ent, eta0, aspect=aspect)
    return time.perf_counter() - t, o["bound"], o["eta"], o["doc_bound"], o["repair"], o["status"], o["nit"]


def slice_csr(ptr, ids, cnt, sel):
    lens = ptr[sel + 1] - ptr[sel]
    p = np.zeros(len(sel) + 1, dtype=np.int64)
    np.cumsum(lens, out=p[1:])
    idx = np.concatenate([np.arange(ptr[d], ptr[d + 1]) for d in sel]) if len(sel) else np.zeros(0, np.int64)
    return p, ids[idx], cnt[idx].astype(np.float64)


class PortPool:
    """The NumPy/SciPy port over `nproc` forked worker processes, created ONCE (outside any timed region)."""

    def __init__(self, nproc):
        import multiprocessing as mp
        self.nproc = nproc
        self.pool = mp.get_context("fork").Pool(nproc)
        self.pool.map(abs, range(nproc))   # workers are up before anything is timed

    def jobs(self, ptr, ids, cnt, beta, mu, siginv, ent, eta0, sel, aspect=None):
        out = []
        for c in (c for c in np.array_split(np.asarray(sel), self.nproc) if len(c)):
            p, i, w = slice_csr(ptr, ids, cnt, c)
            out.append((p, i, w, beta, mu[c], siginv, ent, eta0[c], None if aspect is None else aspect[c]))
        return out

    def run(self, jobs):
        """-> (wall seconds, dict of concatenated per-document results)"""
        t = time.perf_counter()
        res = self.pool.map(_ref_worker, jobs)
        wall = time.perf_counter() - t
        cat = lambda i: np.concatenate([r[i] for r in res], axis=0)  # noqa: E731
        return wall, dict(bound=float(sum(r[1] for r in res)), eta=cat(2), doc_bound=cat(3), repair=cat(4),
                          status=cat(5), nit=cat(6))

    def close(self):
        self.pool.close()
        self.pool.join()


def host_prologue(sigma):
    """stm.py:499-501, the part a reference-side caller keeps: plain NumPy"""
    chol = np.linalg.cholesky(sigma)
    ent = float(np.sum(np.log(np.diag(chol))))
    inv_chol = np.linalg.inv(chol)
    return inv_chol.T * inv_chol, ent


def one_core_as_shipped(ptr, ids, cnt, beta, mu, siginv, ent, eta0, sel, aspect=None, budget_s=6.0):
    """The reference as shipped runs its document loop serially in one process (stm.py:519): time that on a few
    documents of the sample, in THIS process."""
    from oracle import stm_numpy
    n, t0 = 0, time.perf_counter()
    for d in sel:
        p, i, w = slice_csr(ptr, ids, cnt, np.array([d]))
        stm_numpy.estep(p, i, w, beta, mu[[d]], siginv, ent, eta0[[d]], aspect=None if aspect is None else aspect[[d]])
        n += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return {"value": n / (time.perf_counter() - t0), "unit": UNIT, "cores": 1, "docs": n}


def cpu_snapshot_state(args, ptr, ids, cnt, X, aspect, K, V, A, nthreads):
    """The benchmark state (what the first timed E-step of the CUDA arm starts from) prepared on the HOST: spectral
    init (oracle/spectral_numpy.py, sparse Gram) or the reference's random init, then `warmup` EM iterations of the C
    oracle (beta rounded to fp32 after each M-step, as the device stores it)."""
    from oracle import c_oracle, spectral_numpy, stm_numpy
    if args.init == "spectral":
        beta0 = spectral_numpy.spectral_init_fast(ptr, ids, cnt.astype(np.float64), K, V, maxV=5000)[0]
    else:
        beta0 = random_beta(K, V)
    if A > 1:
        beta0 = np.repeat(beta0[None], A, axis=0)
    run = lambda *a, **k: c_oracle.estep(*a, nthreads=nthreads, **k)  # noqa: E731
    if args.warmup == 0:
        D = len(ptr) - 1
        return dict(beta=beta0.astype(np.float32).astype(np.float64), mu=np.zeros((D, K - 1)),
                    sigma=np.eye(K - 1) * 20.0, eta0=np.zeros((D, K - 1)))
    r = stm_numpy.em(ptr, ids, cnt.astype(np.float64), beta0, X, n_iter=args.warmup, threshold=0.0, aspect=aspect,
                     estep_fn=run, round_beta32=True)
    return dict(beta=r["beta"], mu=r["mu"], sigma=r["sigma"], eta0=r["eta"])


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    nproc = os.cpu_count() or 1
    K, V, D, A = workload(args, world)
    t_prep = time.perf_counter()
    ptr, ids, cnt, X = make_corpus(D, V, K, seed=args.seed)
    aspect = (np.arange(D) % A).astype(np.int32) if A > 1 else None
    snap = cpu_snapshot_state(args, ptr, ids, cnt, X, aspect, K, V, A, nproc)
    siginv, ent = host_prologue(snap["sigma"])
    n_s = min(D, args.ref_docs_per_core * nproc)
    sel = np.sort(np.random.default_rng(5).choice(D, size=n_s, replace=False))
    pool = PortPool(nproc)
    jobs = pool.jobs(ptr, ids, cnt, snap["beta"], snap["mu"], siginv, ent, snap["eta0"], sel, aspect)
    small = pool.jobs(ptr, ids, cnt, snap["beta"], snap["mu"], siginv, ent, snap["eta0"], sel[:nproc * 2], aspect)
    t_prep = time.perf_counter() - t_prep
    for _ in range(args.warmup):
        pool.run(small)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.run(jobs)
    wall = time.perf_counter() - t0
    one = one_core_as_shipped(ptr, ids, cnt, snap["beta"], snap["mu"], siginv, ent, snap["eta0"], sel, aspect)
    pool.close()
    value = n_s * args.steps / wall
    sample = (f"{n_s} random documents/step ({args.ref_docs_per_core} per core) of the same corpus and the same state as "
              f"the CUDA arm's first timed E-step ({args.init} init + {args.warmup} EM iterations, prepared on the host in "
              f"{t_prep:.0f} s, untimed); NumPy/SciPy E-step (oracle/stm_numpy.py == reference arithmetic) in {nproc} "
              f"processes forked before the timed region")
    if A > 1:
        sample += ("; config 5: the reference's own kappa update (mnreg) cannot run (SURVEY 8a), so the host-side warm-up "
                   "iterations use its LDA-style update_beta — the E-step timed is the same function on a state of the same kind")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nproc, "kind": "port", "sample": sample,
                         "one_core_as_shipped": one},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
class EmRunner:
    """EM iterations of one STM model with per-phase CUDA events (kernel pair / moments + all-reduce / M-step)."""

    def __init__(self, model, torch, dist, world, lib):
        self.m, self.torch, self.dist, self.world, self.lib = model, torch, dist, world, lib
        self.kernel_ms, self.phase_ev, self.sync_s = [], [], []

    def step(self, timed=False):
        m, torch = self.m, self.torch
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timed else None
        if timed:
            ev[0].record()
        m._estep_device()
        if timed:
            ev[1].record()
        L, h, st = self.lib.load(), m._ctx.handle, m._stream()
        self.lib.check(h, L.stm_moments(h, m._ptr("eta"), m._ptr("x"), m._p, m._ptr("stats"), st))
        m._allreduce(m._d["stats"])
        if timed:
            ev[2].record()
        t = time.perf_counter()
        bound = float(m._d["stats"][m._off[2]].item())   # the one host sync of an EM iteration (convergence test)
        if timed:
            self.sync_s.append(time.perf_counter() - t)
            self.kernel_ms.append(m._ctx.estep_kernel_ms())
        m._mstep_device()
        if timed:
            ev[3].record()
            self.phase_ev.append(ev)
        return bound

    def breakdown(self):
        e = self.phase_ev
        f = lambda i, j: float(np.mean([a[i].elapsed_time(a[j]) for a in e]))  # noqa: E731
        return {"estep_call_ms": f(0, 1), "kernel_bfgs_ms": float(np.mean([k[0] for k in self.kernel_ms])),
                "kernel_post_ms": float(np.mean([k[1] for k in self.kernel_ms])),
                "moments_allreduce_ms": f(1, 2), "mstep_ms": f(2, 3),
                "host_wait_for_elbo_ms": 1e3 * float(np.mean(self.sync_s))}


def timed_run(runner, steps, torch, dist, world, dev, sampler=None):
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    if sampler:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = runner.m._ctx.launch_count()
    w0 = time.perf_counter()
    t0.record()
    bounds = [runner.step(timed=True) for _ in range(steps)]
    t1.record()
    barrier()
    wall = time.perf_counter() - w0
    clocks = sampler.stop() if sampler else None
    tt = torch.tensor([t0.elapsed_time(t1), 1e3 * wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return float(tt[0]), float(tt[1]), bounds, runner.m._ctx.launch_count() - l0, clocks


def fp64_peak_tflops(torch, dev):
    """measured fp64 peak of this GPU: cuBLAS DGEMM 4096^3 (the fp64 pipe's rate; roofline.fp64 denominator)"""
    n = 4096
    a = torch.randn((n, n), dtype=torch.float64, device=dev)
    b = torch.randn((n, n), dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def ours(args):
    import torch
    import torch.distributed as dist
    from strutopy_b200 import STM, _lib
    if args.tune:
        _lib.DEFAULT_TUNE = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.tune.split(",")}

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    K, V, D, A = workload(args, world)
    if args.scaling == "strong":
        D = (D + world - 1) // world
    ptr, ids, cnt, X = make_corpus(D, V, K, seed=args.seed)
    mean_nd = float(ptr[-1]) / D
    aspect = (np.arange(D) % A).astype(np.int32) if A > 1 else None

    def build(ptr_, ids_, cnt_, X_, asp_):
        t = time.perf_counter()
        kw = dict(init_type=args.init, model_type="STM", device=local_rank, distributed=(world > 1), presharded=True)
        if A > 1:   # BASELINE config 5: content covariate, kappa update on (mnreg), two aspects
            kw.update(A=A, beta_index=asp_, lda_beta=False)
        m = STM((ptr_, ids_, cnt_), range(V), A > 1, K, X_, A > 1, 10 ** 9, 0, 0.0, **kw)
        torch.cuda.synchronize()
        if args.init == "random":
            b = random_beta(K, V)
            m.beta = np.repeat(b[None], A, axis=0) if A > 1 else b
        return m, time.perf_counter() - t

    # BASELINE config 3 names spectral initialisation (the reference's default init_type): it runs on the
    # device inside the constructor (stm_spectral_gram / stm_spectral_finish), outside the timed region
    model, t_init = build(ptr, ids, cnt, X, aspect)
    L, h = _lib.load(), model._ctx.handle
    runner = EmRunner(model, torch, dist, world, _lib)
    # L2 hygiene: the per-step working set (eta, mu, theta, beta_ss, corpus: > 250 MB) exceeds the 126 MB L2
    bounds = [runner.step() for _ in range(args.warmup)]
    # snapshot for the parity / e2e / cpu legs: the state the first timed E-step starts from
    snap = dict(beta=np.array(model.beta), mu=model._d["mu"].cpu().numpy(), sigma=np.array(model.sigma),
                eta0=model._d["eta"].cpu().numpy())
    ms_total, wall_ms, b2, launches, clocks = timed_run(runner, args.steps, torch, dist, world, dev,
                                                         ClockSampler(local_rank))
    bounds += b2
    brk = runner.breakdown()
    tt = torch.tensor([brk["kernel_bfgs_ms"], brk["kernel_post_ms"], brk["estep_call_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    bfgs_ms_mean, post_ms_mean, estep_ms_mean = (float(x) for x in tt)
    docs_total = D * world
    value = docs_total * args.steps / (ms_total * 1e-3)

    # every rank must hold the same replicated model after the M-step (beta, Sigma, gamma)
    replicas_identical = None
    if world > 1:
        chk = torch.cat([model._d["beta64_t"].reshape(-1)[::97], model._d["sigma"].reshape(-1),
                         model._d["gamma_t"].reshape(-1)])
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        replicas_identical = bool(torch.equal(lo_, hi_))

    # ---- e2e: the reference-facing host call (stm_estep_host) with pinned fp64 host buffers ----------
    K1 = K - 1
    siginv, ent = host_prologue(snap["sigma"])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    hb = dict(beta=pin(snap["beta"]), mu=pin(snap["mu"]), siginv=pin(siginv),
              theta=torch.empty((D, K), dtype=torch.float64).pin_memory(),
              bss=torch.empty((A, K, V), dtype=torch.float64).pin_memory(),
              sss=torch.empty((K1, K1), dtype=torch.float64).pin_memory(),
              bound=torch.zeros(1, dtype=torch.float64).pin_memory())
    diag = dict(doc_bound=torch.empty(D, dtype=torch.float64).pin_memory(),
                status=torch.empty(D, dtype=torch.int32).pin_memory(), nit=torch.empty(D, dtype=torch.int32).pin_memory(),
                repair=torch.empty(D, dtype=torch.int32).pin_memory())
    vp = lambda t: t.data_ptr()  # noqa: E731
    # eta is an in/out argument (warm start in, result out): every call gets its own pinned copy of the snapshot's eta,
    # prepared BEFORE the timed region, so that the timed region holds exactly the call a user makes
    n_warm_e2e = min(args.warmup, 3)
    eta_bufs = [pin(snap["eta0"].copy()) for _ in range(n_warm_e2e + args.steps + 1)]

    def host_call(i, with_diag=False):
        d = diag if with_diag else {}
        _lib.check(h, L.stm_estep_host(h, vp(hb["beta"]), vp(hb["mu"]), vp(hb["siginv"]), float(ent), vp(eta_bufs[i]),
                                       vp(hb["theta"]), vp(hb["bss"]), vp(hb["sss"]), vp(hb["bound"]),
                                       vp(d["doc_bound"]) if d else None, vp(d["status"]) if d else None,
                                       vp(d["nit"]) if d else None, vp(d["repair"]) if d else None))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(n_warm_e2e):
        host_call(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        host_call(n_warm_e2e + i)
    e2e_s = time.perf_counter() - t0     # the call blocks until every output is home
    barrier()
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = docs_total * args.steps / float(te[0])
    h2d = 8 * (A * K * V + 2 * D * K1 + K1 + 1)
    d2h = 8 * (D * K1 + D * K + A * K * V + K1 * K1 + 1)
    host_call(n_warm_e2e + args.steps, with_diag=True)   # untimed: the per-document outputs of the snapshot state (parity leg)
    gpu = dict(bound=float(hb["bound"][0]), eta=eta_bufs[-1].numpy().copy(), doc_bound=diag["doc_bound"].numpy().copy(),
               status=diag["status"].numpy().copy(), nit=diag["nit"].numpy().copy(), repair=diag["repair"].numpy().copy())

    # ---- BASELINE config 4 in the same run: the SAME number of documents (100k) split over the N GPUs --------
    strong = None
    if world > 1 and args.scaling == "weak" and not args.no_strong:
        Ds = max(1, D // world)
        sl = slice(0, Ds)
        del model, runner
        torch.cuda.empty_cache()
        m2, t2 = build(ptr[:Ds + 1].copy(), ids[:ptr[Ds]], cnt[:ptr[Ds]], X[sl], None if aspect is None else aspect[sl])
        r2 = EmRunner(m2, torch, dist, world, _lib)
        for _ in range(args.warmup):
            r2.step()
        ms2, wall2, _, _, _ = timed_run(r2, args.steps, torch, dist, world, dev)
        b = r2.breakdown()
        tb = torch.tensor([b[k] for k in sorted(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        b = {k: float(v) for k, v in zip(sorted(b), tb)}
        strong = {"docs_total": Ds * world, "docs_per_gpu": Ds, "ms_per_step": ms2 / args.steps,
                  "value": Ds * world * args.steps / (ms2 * 1e-3), "unit": UNIT,
                  "host_wall_ms_per_step": wall2 / args.steps, "breakdown_max_over_ranks_ms": b,
                  "constructor_s_incl_init": t2,
                  "note": "BASELINE config 4: the documents of ONE GPU's weak-scaling shard split over all GPUs; "
                          "compare with the N=1 run's value for the strong-scaling speed-up"}

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernel (the E-step kernel) --------------------------------------
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        b_doc = mean_nd * (8 + 8 * K) + 16 * K - 4            # SURVEY.md §8d algorithmic bytes / document
        # one E-step = kernel A (BFGS) + kernel B (post-optimisation): the algorithmic bytes are those of the
        # pair, so is the time (their two launches, CUDA events inside stm_estep on the launching stream)
        pair_ms = bfgs_ms_mean + post_ms_mean
        achieved = b_doc * D / (pair_ms * 1e-3) / 1e9   # GB/s, per E-step (this rank's D documents)
        prof = {}
        try:
            with open(os.path.join(ROOT, "profiles", "estep_dram_traffic.json")) as f:
                prof = json.load(f)
        except Exception:
            pass
        # fp64 roofline (HBM is knowingly the wrong roof for this path): executed fp64 flops per
        # document from the ncu instruction counts of the same workload (profiles/), over the live kernel time,
        # against the fp64 rate cuBLAS DGEMM reaches on this GPU now
        fp64 = None
        fpd = prof.get("fp64_flops_per_doc") if args.config == "c3" else None
        if fpd:
            pk = fp64_peak_tflops(torch, dev)
            ach = fpd * D / (pair_ms * 1e-3) / 1e12
            fp64 = {"achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk, "flops_per_doc": fpd,
                    "source": "executed DFMA/DADD/DMUL/DMMA thread instructions of the kernel pair (ncu at EM iteration 1 of "
                              "this workload, profiles/estep_dram_traffic.json) / live kernel time; peak = cuBLAS DGEMM "
                              "4096^3 measured now"}
        # ---- CPU baseline + parity on a bounded sample ---------------------------------------------------
        cpu = None
        parity = None
        if not args.no_cpu:
            from oracle import c_oracle
            nproc = os.cpu_count() or 1
            beta_r = snap["beta"].astype(np.float32).astype(np.float64)    # what the device kernels read
            if world == 1:
                n_s = min(D, args.ref_docs_per_core * nproc)
                sel = np.sort(np.random.default_rng(5).choice(D, size=n_s, replace=False))
                pool = PortPool(nproc)
                jobs = pool.jobs(ptr, ids, cnt, beta_r, snap["mu"], siginv, ent, snap["eta0"], sel, aspect)
                pool.run(pool.jobs(ptr, ids, cnt, beta_r, snap["mu"], siginv, ent, snap["eta0"], sel[:nproc * 2], aspect))
                wall, port = pool.run(jobs)
                one = one_core_as_shipped(ptr, ids, cnt, beta_r, snap["mu"], siginv, ent, snap["eta0"], sel, aspect)
                pool.close()
                cpu = {"value": n_s / wall, "unit": UNIT, "cores": nproc, "kind": "port",
                       "sample": f"{n_s} random documents of the same corpus and state as the first timed E-step; "
                                 f"NumPy/SciPy E-step (oracle/stm_numpy.py == reference arithmetic) in {nproc} processes "
                                 f"forked before the timed region",
                       "one_core_as_shipped": one}
                o = c_oracle.estep(ptr, ids, cnt, beta_r, snap["mu"], siginv, ent, snap["eta0"], aspect=aspect,
                                   nthreads=nproc)
                parity = {
                    "elbo_rel_err_vs_c_oracle_full": abs(gpu["bound"] - o["bound"]) / abs(o["bound"]),
                    "max_abs_eta_err_vs_c_oracle_full": float(np.abs(gpu["eta"] - o["eta"]).max()),
                    "status_nit_repair_mismatches_vs_c_oracle_full": [
                        int((gpu[k] != o[k]).sum()) for k in ("status", "nit", "repair")],
                    # the NumPy port tests positive definiteness with np.linalg.eigvals like the reference
                    # (stm.py:1017), not with the pivot shortcut the C oracle and kernel B share
                    "max_abs_eta_err_vs_numpy_port_sample": float(np.abs(gpu["eta"][sel] - port["eta"]).max()),
                    "max_rel_doc_bound_err_vs_numpy_port_sample": float(np.max(
                        np.abs(gpu["doc_bound"][sel] - port["doc_bound"]) / np.abs(port["doc_bound"]))),
                    "elbo_rel_err_vs_numpy_port_sample": abs(float(gpu["doc_bound"][sel].sum()) - port["bound"]) / abs(port["bound"]),
                    "repair_status_nit_mismatches_vs_numpy_port_sample": [
                        int((gpu[k][sel] != port[k]).sum()) for k in ("repair", "status", "nit")],
                    "repair_rate_sample": float(np.mean(port["repair"] > 0)),
                }
            else:
                # N > 1: rank 0 checks a sample of ITS shard against the C oracle (same snapshot state)
                n_s = min(D, 2000)
                sel = np.sort(np.random.default_rng(5).choice(D, size=n_s, replace=False))
                p_, i_, w_ = slice_csr(ptr, ids, cnt, sel)
                o = c_oracle.estep(p_, i_, w_, beta_r, snap["mu"][sel], siginv, ent, snap["eta0"][sel],
                                   aspect=None if aspect is None else aspect[sel], nthreads=nproc)
                parity = {
                    "rank0_sample_docs": n_s,
                    "elbo_rel_err_vs_c_oracle_sample": abs(float(gpu["doc_bound"][sel].sum()) - o["bound"]) / abs(o["bound"]),
                    "max_abs_eta_err_vs_c_oracle_sample": float(np.abs(gpu["eta"][sel] - o["eta"]).max()),
                    "status_nit_repair_mismatches_vs_c_oracle_sample": [
                        int((gpu[k][sel] != o[k]).sum()) for k in ("status", "nit", "repair")],
                    "replicas_identical_after_mstep": replicas_identical,
                }
        cfg = config_dict(args, world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "mean_distinct_words_per_doc": mean_nd, "constructor_s_incl_init": t_init,
            "host_wall_ms_per_step": wall_ms / args.steps, "breakdown_ms": brk,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": prof.get("dram_bytes_per_launch") if args.config == "c3" else None,
                         "traffic_by_kernel": prof.get("by_kernel") if args.config == "c3" else None,
                         "kernel": "stm::bfgs_kernel + stm::post_group_kernel (one E-step = this launch pair)",
                         "bytes_per_doc": b_doc, "estep_ms_per_launch": pair_ms,
                         "kernel_ms": {"stm::bfgs_kernel": bfgs_ms_mean, "stm::post_group_kernel": post_ms_mean,
                                       "estep_call_incl_memsets_epilogue": estep_ms_mean},
                         "dominant_kernel": "stm::bfgs_kernel" if bfgs_ms_mean >= post_ms_mean else "stm::post_group_kernel",
                         "dominant_kernel_share_of_estep": max(bfgs_ms_mean, post_ms_mean) / pair_ms,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "fp64": fp64,
                         "limiter": "kernel A: instruction fetch (gcc__cache_requests_type_instruction at 97 % of its peak "
                                    "rate, no_instruction the first stall reason); kernel B: fixed-latency dependencies, named "
                                    "barriers and instruction fetch (88 % of that peak) at IPC 1.8 "
                                    "(profiles/r02d_estep_pair_ncu_summary.txt, r02_tuning_log.md)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "stm_estep_host (C ABI, fp64 host buffers in the reference's layouts; corpus resident; "
                           "documents chunked so that copies overlap kernels)",
                    "state": f"every call runs the E-step of the snapshot state (EM iteration {args.warmup}); `value` "
                             f"averages the EM iterations {args.warmup}..{args.warmup + args.steps - 1}, whose E-steps "
                             "differ in cost by a few percent"},
            "gpu_launches": int(launches), "clocks": clocks,
            "elbo_trace_tail": bounds[-3:], "parity": parity, "strong": strong,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c5"],
                    help="c3: BASELINE configs[2]/[3] (the metric's config); c5: configs[4]'s shape")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--docs", type=int, default=0, help="documents per GPU (weak) or in total (strong); 0: the config's")
    ap.add_argument("--K", type=int, default=50)
    ap.add_argument("--V", type=int, default=10000)
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--init", default="spectral", choices=["spectral", "random"],
                    help="beta initialisation (BASELINE config 3: spectral)")
    ap.add_argument("--ref-docs-per-core", type=int, default=48)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline / parity legs")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling block")
    ap.add_argument("--tune", default="", help="stm_tune overrides for every context, e.g. host_chunks=8 (A/B runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
