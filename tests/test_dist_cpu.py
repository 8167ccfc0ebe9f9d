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


def _bn_worker(rank, world, port, split, q):
    sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import mmb_dp
    torch.manual_seed(3)
    B, d = 9, 7
    x_all = torch.randn(B, d) * 2 + 1
    g_all = torch.randn(B, d)
    bn = torch.nn.BatchNorm1d(d)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(d) + 0.5)
        bn.bias.copy_(torch.randn(d))
    lo, hi = (0, split) if rank == 0 else (split, B)
    dp = mmb_dp.DataParallel(None, None)
    out = []
    for step in range(2):                       # two steps: running buffers accumulate
        x = x_all[lo:hi].clone().requires_grad_(True)
        if hi > lo:
            y = mmb_dp.sync_batch_norm(bn, x, dp)
            (y * g_all[lo:hi]).sum().backward()
            gw, gb, gx = bn.weight.grad.clone(), bn.bias.grad.clone(), x.grad.clone()
            bn.weight.grad = bn.bias.grad = None
        else:                                   # this rank owns no member of the batch
            bwd = mmb_dp.sync_batch_norm_empty(bn, dp, torch.device('cpu'))
            gw, gb = bwd()
            y, gx = torch.zeros(0, d), torch.zeros(0, d)
        out.append((y.detach().numpy(), gx.numpy(), gw.numpy(), gb.numpy()))
    q.put((rank, out, bn.running_mean.numpy(), bn.running_var.numpy(), int(bn.num_batches_tracked)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('split', [4, 0, 9])
def test_sync_batch_norm_equals_full_batch(split):
    """Data-parallel MMB step: BatchNorm1d statistics over the GLOBAL batch -- outputs, input
    gradients, affine gradients and running buffers equal torch's BatchNorm1d on the concatenated batch,
    for an uneven split and for a rank that owns no member of the batch."""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bn_worker, args=(r, world, port, split, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(3)
    B, d = 9, 7
    x_all = torch.randn(B, d) * 2 + 1
    g_all = torch.randn(B, d)
    bn = torch.nn.BatchNorm1d(d)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(d) + 0.5)
        bn.bias.copy_(torch.randn(d))
    for step in range(2):
        x = x_all.clone().requires_grad_(True)
        y = bn(x)
        (y * g_all).sum().backward()
        y_dp = np.concatenate([res[0][1][step][0], res[1][1][step][0]])
        gx_dp = np.concatenate([res[0][1][step][1], res[1][1][step][1]])
        np.testing.assert_allclose(y_dp, y.detach().numpy(), rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(gx_dp, x.grad.numpy(), rtol=1e-4, atol=1e-5)
        for r in range(world):                                   # affine gradients: global on EVERY rank
            np.testing.assert_allclose(res[r][1][step][2], bn.weight.grad.numpy(), rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(res[r][1][step][3], bn.bias.grad.numpy(), rtol=1e-4, atol=1e-5)
        bn.weight.grad = bn.bias.grad = None
    for r in range(world):
        np.testing.assert_allclose(res[r][2], bn.running_mean.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(res[r][3], bn.running_var.numpy(), rtol=1e-5, atol=1e-6)
        assert res[r][4] == 2


def test_global_batches_partition():
    """Every rank draws the same global index batches (same torch seed) and the ranks' members partition
    each batch; an all-reduced sum of local_sum / B_global equals the global mean."""
    sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
    import mmb_dp
    import sif_dist
    from torch.utils.data import DataLoader
    n, world = 23, 3
    vals = torch.arange(n, dtype=torch.float64) ** 2
    per_rank = []
    for rank in range(world):
        torch.manual_seed(11)
        loader = DataLoader(range(n), batch_size=5, shuffle=True)
        per_rank.append([mmb_dp.global_index_batches(loader) for _ in range(2)])
    torch.manual_seed(11)
    ref_loader = DataLoader(range(n), batch_size=5, shuffle=True)
    want = [[list(map(int, b)) for b in ref_loader] for _ in range(2)]      # what iterating the loader yields
    assert per_rank[0] == per_rank[1] == per_rank[2] == want
    for batch in want[0]:
        parts = []
        for rank in range(world):
            lo, hi = sif_dist.shard_bounds(n, world, rank)
            parts.append([b for b in batch if lo <= b < hi])
        assert sorted(sum(parts, [])) == sorted(batch)
        total = sum(float(vals[p].sum()) / len(batch) for p in parts if p)
        assert abs(total - float(vals[batch].mean())) < 1e-9
