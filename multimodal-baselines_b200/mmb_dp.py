"""Data-parallel MMB latent optimisation over the GPUs of one box (SURVEY.md §8e "MMB training").

The reference has no distributed code; this is the one strategy its loop admits (reference
simplesif.py:49-162, e2e loop 708-790):

  * utterances -- data rows AND their latent rows, which are per-utterance parameters -- shard in the
    same contiguous blocks as the SIF stage (``sif_dist.shard_bounds``): latents never travel;
  * the generator heads (843,400 parameters for MMB2) are replicated; every step their gradients are
    summed over the ranks with ONE exchange over NVLink peer memory (``PeerComm.allreduce_`` on the flat
    gradient buffer ``HeadsFunction.backward`` writes into -- the 4 MiB exchange slot holds it), issued
    right behind the kernel that produces them;
  * the objective is the mean over the GLOBAL batch (reference simplesif.py:133 ``log_prob.mean()``):
    each rank scales the sum over ITS members of the batch by 1 / B_global, so the all-reduced
    gradients are exactly the single-process ones;
  * ``norm='batch_norm'`` (half of the reference grid, models.py:161-168): batch statistics are taken
    over the global batch (``SyncBatchNormFunction``: one small exchange forward, one backward), and the
    running buffers are updated with them on every rank; ``layer_norm`` / ``None`` are per row and need
    nothing (their affine gradients ride in the step's tail exchange together with the loss value).

Every rank draws the SAME global index batches (the DataLoader's sampler under the same torch seed) and
takes the members it owns; a rank that owns none of a batch still joins every exchange of the step with
zeros, in the same order.  All sums are taken in rank order (identical bits on every rank), so the
replicated heads stay bit-identical without a broadcast.

The exchange goes through ``PeerComm`` on NCCL process groups of one box and through
``torch.distributed.all_reduce`` otherwise (gloo on the CPU: tests/test_dist_cpu.py checks the partition
arithmetic and the synchronised BatchNorm there).
"""
import time

import numpy as np
import torch
import torch.distributed as dist



class DataParallel(object):
    """What the step's kernels need to know about the job: the exchange and the global batch size."""

    def __init__(self, comm=None, group=None):
        self.comm = comm               # sif_dist.PeerComm or None (-> torch.distributed)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.batch_global = None       # set per step
        self.exchanges = 0

    def allreduce_(self, t):
        """In-place sum over ranks, rank order.  float32 / float64, contiguous."""
        self.exchanges += 1
        if self.world == 1:
            return t
        if self.comm is not None and t.is_cuda:
            return self.comm.allreduce_(t)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


_ACTIVE = None


def active():
    """The DataParallel context of the running step (None outside ``optimize_latents_dp``)."""
    return _ACTIVE


class _Scope(object):
    def __init__(self, dp):
        self.dp = dp

    def __enter__(self):
        global _ACTIVE
        self.prev, _ACTIVE = _ACTIVE, self.dp
        return self.dp

    def __exit__(self, *exc):
        global _ACTIVE
        _ACTIVE = self.prev


class SyncBatchNormFunction(torch.autograd.Function):
    """``nn.BatchNorm1d`` in training mode (reference models.py:164, never put in ``.eval()``:
    SURVEY.md §3.4) with the statistics of the GLOBAL batch: sum and sum of squares of the local rows are
    exchanged (2 d + 1 numbers, float64), every rank normalises its rows with the same mean / variance
    and updates its running buffers with them.  Backward exchanges sum(dy) and sum(dy * xhat): they are
    at once the affine gradients (already global -- not reduced again) and the correction of dx."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, num_batches_tracked, momentum, eps, dp):
        B_local, d = x.shape
        stats = torch.zeros(2 * d + 1, dtype=torch.float64, device=x.device)
        if B_local > 0:
            xd = x.double()
            stats[:d] = xd.sum(0)
            stats[d:2 * d] = (xd * xd).sum(0)
            stats[2 * d] = B_local
        dp.allreduce_(stats)
        n = stats[2 * d]
        mean = stats[:d] / n
        var = (stats[d:2 * d] / n - mean * mean).clamp_min(0.)       # biased, as F.batch_norm normalises with
        invstd = torch.rsqrt(var + eps)
        with torch.no_grad():
            if running_mean is not None:
                running_mean.mul_(1 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
                unbiased = var * (n / (n - 1).clamp_min(1.))
                running_var.mul_(1 - momentum).add_(unbiased.to(running_var.dtype), alpha=momentum)
            if num_batches_tracked is not None:
                num_batches_tracked += 1
        mean32, invstd32 = mean.float(), invstd.float()
        xhat = (x - mean32) * invstd32
        ctx.save_for_backward(xhat, weight, invstd32)
        ctx.dp, ctx.n = dp, n
        return xhat * weight + bias

    @staticmethod
    def backward(ctx, dy):
        xhat, weight, invstd = ctx.saved_tensors
        d = xhat.shape[1]
        sums = torch.zeros(2 * d, dtype=torch.float64, device=dy.device)
        if xhat.shape[0] > 0:
            dyd = dy.double()
            sums[:d] = dyd.sum(0)
            sums[d:] = (dyd * xhat.double()).sum(0)
        ctx.dp.allreduce_(sums)
        n = ctx.n
        dbias = sums[:d].float()
        dweight = sums[d:].float()
        dx = (dy - (sums[:d] / n).float() - xhat * (sums[d:] / n).float()) * (weight * invstd)
        return dx, dweight, dbias, None, None, None, None, None, None


def sync_batch_norm(bn, x, dp):
    """``bn(x)`` for an ``nn.BatchNorm1d`` in training mode with global-batch statistics."""
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return SyncBatchNormFunction.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                       bn.num_batches_tracked, momentum, bn.eps, dp)


def sync_batch_norm_empty(bn, dp, device):
    """A rank with no member of the batch: the two exchanges of SyncBatchNormFunction with zeros (forward
    now; the caller invokes the returned thunk where backward would run), buffers updated like everyone's."""
    d = bn.num_features
    x = torch.zeros((0, d), dtype=torch.float32, device=device)
    w = bn.weight.detach().requires_grad_(True)
    b = bn.bias.detach().requires_grad_(True)
    y = sync_batch_norm_with(x, w, b, bn, dp)

    def backward():
        y.sum().backward()
        return w.grad, b.grad
    return backward


def sync_batch_norm_with(x, w, b, bn, dp):
    momentum = 0.1 if bn.momentum is None else bn.momentum
    return SyncBatchNormFunction.apply(x, w, b, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                                       momentum, bn.eps, dp)


def global_index_batches(index_loader):
    """The index batches a DataLoader over the GLOBAL utterance range produces this epoch, as host lists,
    with the same draws from torch's generator as iterating it (num_workers = 0: one base-seed draw when
    the iterator is built, then the sampler's own) -- every rank calls this under the same seed."""
    torch.empty((), dtype=torch.int64).random_(generator=index_loader.generator)
    return [list(b) for b in index_loader.batch_sampler]


def optimize_latents_dp(args, train, gen_model, embed_local, local_dataset, index_loader, n_epochs, lr,
                        word_prob_fn, device, lo, comm='auto', group=None, verbose=True):
    """``simplesif.optimize_latents`` (reference simplesif.py:49-162) with the utterances sharded over the
    ranks of ``group``.

    ``embed_local`` (n_local, d) and ``local_dataset`` (an ``MMData`` over rows [lo, lo + n_local) of the
    split) are this rank's block; ``index_loader`` is a DataLoader over ``range(n_global)`` with the
    reference's batch size / shuffle flag, used only for its sampler.  ``gen_model`` must hold the same
    parameters on every rank.  Returns ``(local latents (n_local, d) on device, (losses, []))`` with
    ``losses`` the global epoch sums, identical on every rank."""
    import simplesif
    import sif_dist
    from losses import get_log_prob_matrix
    if comm == 'auto':
        comm = sif_dist.default_comm(group)
    dp = DataParallel(comm, group)
    embeddings = torch.tensor(np.array(embed_local, copy=True), device=device, dtype=torch.float32)
    embeddings.requires_grad = True
    n_local = embeddings.shape[0]
    hi = lo + n_local
    train_heads = bool(train and not args['freeze_weights'])
    grad_params = [embeddings]
    if train_heads:
        grad_params.extend(gen_model.parameters())
    optimizer = simplesif._make_optimizer(args, grad_params, lr)
    moments = simplesif._dataset_moments(args, local_dataset) if n_local > 0 else None
    norm = gen_model.norm
    is_bn = isinstance(norm, torch.nn.BatchNorm1d)
    norm_params = [p for p in (norm.parameters() if norm is not None else []) if p.requires_grad]
    head_params = [p for p in gen_model.embed2out.parameters()]
    tail_n = (sum(p.numel() for p in norm_params) if (train_heads and not is_bn) else 0) + 1

    losses = []
    start_time = time.time()
    with _Scope(dp):
        dp.reduce_head_grads = train_heads
        for i in range(n_epochs):
            epoch_loss = torch.zeros((), dtype=torch.float32, device=device)
            iters = 0
            for batch in global_index_batches(index_loader):
                iters += 1
                dp.batch_global = len(batch)
                mine = [b - lo for b in batch if lo <= b < hi]
                optimizer.zero_grad()
                tail = torch.zeros(tail_n, dtype=torch.float32, device=device)
                if mine:
                    j = torch.tensor(mine, dtype=torch.int64, device=device)
                    x = local_dataset[j]
                    _, batch_data, batch_masks = simplesif._batch_dicts(args, simplesif._with_moments(args, x, moments),
                                                                        getattr(local_dataset, 'table', None))
                    e = embeddings[j]
                    out = gen_model(e)              # BatchNorm: global statistics; heads: gradients all-reduced
                    log_prob = -get_log_prob_matrix(args, e, out, batch_data, batch_masks, word_prob_fn,
                                                    device=device, verbose=False)
                    loss = log_prob.sum() / dp.batch_global
                    loss.backward()
                    tail[-1] = loss.detach()
                else:
                    # no member of this batch lives here: join the step's exchanges with zeros, in the order
                    # the other ranks issue them (BN forward, heads gradient, BN backward, tail)
                    bn_backward = sync_batch_norm_empty(norm, dp, device) if is_bn else None
                    if train_heads:
                        flat = torch.zeros(sum(p.numel() for p in head_params), dtype=torch.float32, device=device)
                        dp.allreduce_(flat)
                        off = 0
                        for p in head_params:
                            p.grad = flat[off:off + p.numel()].view_as(p)
                            off += p.numel()
                    if bn_backward is not None:
                        gw, gb = bn_backward()
                        if train_heads:
                            norm.weight.grad, norm.bias.grad = gw, gb
                    embeddings.grad = torch.zeros_like(embeddings)      # Adam still moves every row (stale moments)
                # tail: LayerNorm affine gradients (local sums) + the loss value
                if tail_n > 1:
                    off = 0
                    for p in norm_params:
                        if p.grad is not None:
                            tail[off:off + p.numel()] = p.grad.reshape(-1)
                        off += p.numel()
                dp.allreduce_(tail)
                if tail_n > 1:
                    off = 0
                    for p in norm_params:
                        p.grad = tail[off:off + p.numel()].view_as(p).clone()
                        off += p.numel()
                optimizer.step()
                epoch_loss += tail[-1]
            losses.append(float(epoch_loss))
            if verbose and i % 10 == 0:
                print("epoch {}: {} ({}s)".format(i, losses[-1] / max(iters, 1), time.time() - start_time))
    if comm is not None:
        comm.check()
    embeddings.requires_grad = False
    return embeddings, (losses, [])
