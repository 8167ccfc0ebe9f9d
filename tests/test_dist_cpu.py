"""CPU (gloo, world_size 2): the partition / exchange arithmetic of the sharded SIF path.
The kernels need a GPU; what is checked here is the host logic around them -- shard bounds,
the all-reduce of the d x d Gram and of the N < d start block, and that the replicated
solve sees the same inputs on every rank."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_global, d, q):
    sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sif_dist
    from oracle import sif_oracle as so
    rng = np.random.default_rng(7)
    X = rng.standard_normal((n_global, d)).astype(np.float32)           # same on every rank
    lo, hi = sif_dist.shard_bounds(n_global, world, rank)
    Xl = torch.as_tensor(X[lo:hi])
    G = (Xl.T @ Xl).float()
    sif_dist.allreduce_sum_(G)
    # the rank's share of S0 = X^T Omega: its rows of X against ITS rows of the global seeded Omega (the product
    # itself is mmb_start_block_xt on the GPU; here the partition rule is what is under test)
    S0 = Xl.double().T @ torch.as_tensor(sif_dist.local_omega_rows(n_global, lo, hi - lo, 1, so.start_block))
    sif_dist.allreduce_sum_(S0)
    q.put((rank, lo, hi, G.numpy(), S0.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_global', [37, 600])
def test_gram_allreduce_two_ranks(n_global):
    d, world = 48, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_global, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
    import sif_dist
    from oracle import sif_oracle as so
    rng = np.random.default_rng(7)
    X = rng.standard_normal((n_global, d)).astype(np.float32)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n_global     # contiguous cover
    want_G = X.astype(np.float64).T @ X.astype(np.float64)
    want_S0 = X.astype(np.float64).T @ so.start_block(n_global, 1)
    for _rank, _lo, _hi, G, S0 in res:
        np.testing.assert_allclose(G, want_G, rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(S0, want_S0, rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(res[0][3], res[1][3])        # every rank solves on the same bits
    np.testing.assert_allclose(sif_dist.cpu_reference_partition([X[:res[0][2]], X[res[0][2]:]]), want_G,
                               rtol=1e-6, atol=1e-6)


def test_shard_bounds_cover():
    sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
    import sif_dist
    for n in (0, 1, 7, 8, 10_000_000):
        for w in (1, 2, 3, 8):
            b = [sif_dist.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
