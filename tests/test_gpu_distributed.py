"""Data-parallel EM over two ranks against the single-process run: NCCL on two GPUs when the box has them; on a
one-GPU box the SAME test runs with both ranks on cuda:0 and the gloo backend (which all-reduces CUDA tensors through
the host), so that the sharded path — nnz-balanced shards, the one all-reduce per iteration, the replicated M-step,
sharded spectral init, eval_heldout, save_model — is exercised wherever the GPU tests run."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
import numpy as np, torch, torch.distributed as dist
from conftest import load_golden
from strutopy_b200 import STM
rank = int(os.environ['RANK'])
two = torch.cuda.device_count() >= 2
dev_index = rank if two else 0
torch.cuda.set_device(dev_index)
if two:
    dist.init_process_group('nccl', device_id=torch.device('cuda', dev_index))
else:
    dist.init_process_group('gloo')
g = load_golden('em_c1.npz'); K, V = int(g['K']), int(g['V'])
docs = (g['doc_ptr'], g['word_id'], g['count'])
def fit(distributed):
    m = STM(docs, range(V), False, K, g['X'], False, 6, 0, 0.0, init_type='random', model_type='STM',
            device=dev_index, distributed=distributed)
    m.beta = g['beta0']
    m.expectation_maximization(saving=False)
    return m
md = fit(True)
ms = fit(False)
assert md.N_local < md.N and md.N == ms.N
rel = np.abs((np.array(md.last_bounds) - np.array(ms.last_bounds)) / np.array(ms.last_bounds))
assert len(md.last_bounds) == 6 and rel.max() < 1e-6, rel          # summation order differs; EM amplifies
assert rel[0] < 1e-12, rel
ref = g['bounds'][:6]
assert np.abs((np.array(md.last_bounds) - ref) / ref).max() < 1e-4
np.testing.assert_allclose(md.theta, ms.theta, atol=1e-4)            # gathered over ranks
np.testing.assert_allclose(md.beta, ms.beta, atol=1e-6)
np.testing.assert_allclose(md.gamma, ms.gamma, atol=1e-5)
np.testing.assert_allclose(md.sigma, ms.sigma, atol=1e-6)
# held-out likelihood and persistence over sharded documents (ADVICE r01): global indexing, one mean, rank 0 writes
held = docs
hd, hs = md.eval_heldout(held), ms.eval_heldout(held)
assert abs(hd - hs) <= 1e-6 * abs(hs), (hd, hs)
out = os.path.join({tmp!r}, 'model')
md.save_model(out)
dist.barrier()
assert np.load(os.path.join(out, 'theta_hat.npy')).shape == (md.N, K)
assert np.load(os.path.join(out, 'X.npy')).shape[0] == md.N
# every rank holds the same replicated model after the M-step
for name in ('beta', 'sigma', 'gamma'):
    t = torch.from_numpy(np.ascontiguousarray(getattr(md, name))).cuda()
    lo_, hi_ = t.clone(), t.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN); dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    assert torch.equal(lo_, hi_), name
bss, sss = md.E_step(); md.M_step(bss, sss)
bs2, ss2 = ms.E_step(); ms.M_step(bs2, ss2)
np.testing.assert_allclose(bss, bs2, atol=1e-5); np.testing.assert_allclose(md.sigma, ms.sigma, atol=1e-6)
# spectral initialisation over sharded documents: local Gram statistics, ONE all-reduce, replicated recovery
sp = load_golden('spectral.npz')
sdocs = (sp['f_doc_ptr'], sp['f_word_id'], sp['f_count'].astype(np.float64))
Ks, Vs = int(sp['f_cfg'][2]), sp['f_beta'].shape[1]
Xs = (np.arange(len(sdocs[0]) - 1) % 2).astype(np.float64)[:, None]
bd = STM(sdocs, range(Vs), False, Ks, Xs, False, 2, 0, 0.0, init_type='spectral', device=dev_index, distributed=True).beta
b1 = STM(sdocs, range(Vs), False, Ks, Xs, False, 2, 0, 0.0, init_type='spectral', device=dev_index, distributed=False).beta
assert np.abs(bd - b1).max() <= 2e-7 * b1.max(), np.abs(bd - b1).max()   # beta is stored in fp32
assert np.abs(bd - sp['f_beta']).max() <= 2e-7 * sp['f_beta'].max()
dist.barrier(); dist.destroy_process_group()
print('rank', rank, 'ok')
"""


def test_two_gpu_em_matches_single_gpu(tmp_path):
    import torch
    assert torch.cuda.is_available()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, tmp=str(tmp_path)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for pr in procs:
        out, _ = pr.communicate(timeout=600)
        assert pr.returncode == 0, out[-4000:]
