"""Downstream regressors of SEVERAL grid points trained at once.

``sentiment_model.train_sentiment_for_latents`` (reference sentiment_model.py:165-265) fits one
300 -> {100, 150} -> n_out MLP per grid point with plain SGD on the mean L1 loss: 400 epochs x 41 batches of 32
x ~25 tiny kernels -- 1.0 s of a 2.2 s grid point on a B200, all of it launch latency.  The grid points of a
sweep are independent and their regressors have the same shape, so K of them run as ONE batched model:
parameters ``(K, H, d)``, every step one ``baddbmm`` per layer over a leading config dimension, per-config
step sizes, whole epochs replayed as one CUDA graph.  Numerics stay torch's (batched GEMMs instead of K
separate ones); the reference module remains the parity yardstick (tests/test_mmb_gpu.py).

What makes the batched run EQUAL to K sequential ones is the random stream: each grid point's regressor draws
its initialisation and its shuffles from torch's global generator right after that point's latent optimisation.
``RegressorJob`` snapshots the generator there; ``record_schedule`` replays exactly the draws
``train_sentiment_for_latents`` would make (model init; one DataLoader base seed + one permutation per pass:
the initial scoring pass, every training epoch, the validation pass every 10 epochs, the final scoring pass)
and keeps the index batches, so the batched trainer visits the same samples in the same order per config.

Only the configuration the reference grid uses is batched: ``early_stopping`` off (make_configs.py has no such
key; simplesif.py's default is off).  Anything else goes through the sequential module.
"""
import numpy as np
import torch
from torch.utils.data import DataLoader

import graph_capture
from sentiment_model import SentimentModel, _score


class RegressorJob(object):
    """One grid point's regressor problem, captured where the sequential code would start it."""

    def __init__(self, args, latents, labels, rng_state=None, tag=None):
        self.args = dict(args)
        self.latents = latents                      # (train, valid, test) device tensors (N_s, d)
        self.labels = labels                        # (train, valid, test) NumPy arrays
        self.rng_state = torch.get_rng_state() if rng_state is None else rng_state
        self.tag = tag
        self.results = None
        self.train_losses = self.valid_losses = None

    def key(self):
        tr, va, te = self.labels
        return (self.args['sentiment_hidden_size'], self.args['n_sentiment_epochs'], tuple(tr.shape), tuple(va.shape),
                tuple(te.shape), tuple(self.latents[0].shape))


def can_batch(args, labels):
    """The batched trainer covers what the reference grid uses: no early stopping, labels (N,) or (N, n_out > 1).
    ((N, 1) labels make the reference's L1 broadcast to (B, B) -- sentiment_model._l1 -- and stay sequential.)"""
    train = np.asarray(labels[0])
    return (not args.get('early_stopping', False)) and (train.ndim == 1 or (train.ndim == 2 and train.shape[-1] > 1))


def _pass_indices(loader):
    """One pass of ``for j, y in loader``: the iterator's base-seed draw, then the sampler's permutation."""
    torch.empty((), dtype=torch.int64).random_(generator=loader.generator)
    return np.fromiter((i for b in loader.batch_sampler for i in b), dtype=np.int64, count=len(loader.dataset))


def record_schedule(job, valid_niter=10):
    """Replay the global-generator draws of ``train_sentiment_for_latents`` for this job (nothing is computed):
    returns the CPU-initialised model and the index order of every pass."""
    torch.set_rng_state(job.rng_state)
    d = job.latents[0].shape[-1]
    train, valid, test = job.labels
    n_out = 1 if train.ndim == 1 else train.shape[-1]
    model = SentimentModel(d, job.args['sentiment_hidden_size'], n_out)          # the init draws
    loaders = [DataLoader(range(len(s)), batch_size=32, shuffle=True) for s in (train, valid, test)]
    before = _pass_indices(loaders[2])                                           # initial scoring pass
    epochs, valids = [], []
    for i in range(job.args['n_sentiment_epochs']):
        epochs.append(_pass_indices(loaders[0]))
        if i % valid_niter == 0:
            valids.append(_pass_indices(loaders[1]))
    after = _pass_indices(loaders[2])                                            # final scoring pass
    job.final_rng_state = torch.get_rng_state()
    return model, dict(before=before, epochs=epochs, valids=valids, after=after)


class _BatchedMLP(object):
    """K regressors with a leading config dimension: parameters (K, H, d), (K, H), (K, n_out, H), (K, n_out)."""

    def __init__(self, models, device):
        st = lambda name: torch.stack([dict(m.named_parameters())[name].detach() for m in models]).to(device)
        self.W1, self.b1 = st('hidden1.weight').requires_grad_(True), st('hidden1.bias').requires_grad_(True)
        self.W2, self.b2 = st('out.weight').requires_grad_(True), st('out.bias').requires_grad_(True)
        self.params = [self.W1, self.b1, self.W2, self.b2]

    def forward(self, x):                           # x (K, b, d) -> (K, b, n_out)
        h = torch.relu(torch.baddbmm(self.b1[:, None, :], x, self.W1.transpose(1, 2)))
        return torch.baddbmm(self.b2[:, None, :], h, self.W2.transpose(1, 2))


def _l1_per_config(y, t):
    """Mean L1 per config as the reference's ``nn.L1Loss(reduce=False)(model(x), t).mean()`` computes it for (B,)
    labels with n_out = 1 and for (B, n_out) labels (the squeezed prediction has the label's shape there)."""
    if t.dim() == 2:                                # labels (K, b): prediction (K, b, 1) squeezed
        y = y.squeeze(-1)
    return (y - t).abs().flatten(1).mean(1)


def train_batched(jobs, device, valid_niter=10, use_graph=True):
    """Train the regressors of ``jobs`` (same ``key()``) together; fills ``job.results`` (the reference's
    ``test_results_after`` metrics), ``job.train_losses``, ``job.valid_losses``."""
    assert len(set(j.key() for j in jobs)) == 1
    K = len(jobs)
    sched, models = [], []
    for job in jobs:
        m, s = record_schedule(job, valid_niter)
        models.append(m)
        sched.append(s)
    net = _BatchedMLP(models, device)
    lr = torch.tensor([j.args['sentiment_lr'] for j in jobs], dtype=torch.float32, device=device)
    f32 = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=device)
    X = [torch.stack([j.latents[s].detach() for j in jobs]) for s in range(3)]          # (K, N_s, d)
    Y = [torch.stack([f32(j.labels[s]) for j in jobs]) for s in range(3)]               # (K, N_s[, n_out])
    ar = torch.arange(K, device=device)[:, None]
    n_train = X[0].shape[1]
    sizes = [min(32, n_train - o) for o in range(0, n_train, 32)]
    n_epochs = jobs[0].args['n_sentiment_epochs']

    static_idx = torch.zeros((K, n_train), dtype=torch.int64, device=device)

    def epoch_body():
        total = torch.zeros(K, device=device)
        off = 0
        for b in sizes:
            j = static_idx[:, off:off + b]
            off += b
            loss = _l1_per_config(net.forward(X[0][ar, j]), Y[0][ar, j])                # (K,)
            grads = torch.autograd.grad(loss.sum(), net.params)
            with torch.no_grad():
                for p, g in zip(net.params, grads):
                    p.sub_(g * lr.view(-1, *([1] * (p.dim() - 1))))                     # SGD, one step size per config
            total = total + loss.detach()
        return total

    graph = static_total = None
    if use_graph and X[0].is_cuda:
        saved = [p.detach().clone() for p in net.params]
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            epoch_body()                                                                 # warm-up (allocator, lazy inits)
        torch.cuda.current_stream(device).wait_stream(side)
        with torch.no_grad():
            for p, s_ in zip(net.params, saved):
                p.copy_(s_)
        graph = torch.cuda.CUDAGraph()
        with graph_capture.capture(graph):
            static_total = epoch_body()

    def eval_pass(split, order):
        """Mean over batches of the batch-mean L1 (what train_sentiment's validation pass averages), per config."""
        idx = torch.as_tensor(np.stack(order)).to(device)                                # (K, N)
        tot, nb = torch.zeros(K, device=device), 0
        with torch.no_grad():
            for o in range(0, idx.shape[1], 32):
                j = idx[:, o:o + 32]
                tot += _l1_per_config(net.forward(X[split][ar, j]), Y[split][ar, j])
                nb += 1
        return tot / max(nb, 1)

    train_sums, valid_vals = [], []
    all_epochs = torch.as_tensor(np.stack([np.stack([s['epochs'][i] for s in sched]) for i in range(n_epochs)]))   # (E, K, N)
    all_epochs = all_epochs.pin_memory() if X[0].is_cuda else all_epochs
    for i in range(n_epochs):
        static_idx.copy_(all_epochs[i], non_blocking=True)
        if graph is not None:
            graph.replay()
            train_sums.append(static_total.clone())
        else:
            train_sums.append(epoch_body())
        if i % valid_niter == 0:
            valid_vals.append(eval_pass(1, [s['valids'][i // valid_niter] for s in sched]))
    train_sums = torch.stack(train_sums).cpu().numpy() / len(sizes)                      # (E, K)
    valid_vals = torch.stack(valid_vals).cpu().numpy() if valid_vals else np.zeros((0, K))
    # final scoring pass: metrics do not depend on the order of the samples
    with torch.no_grad():
        pred = net.forward(X[2])                                                         # (K, N_test, n_out)
    for k, job in enumerate(jobs):
        p = pred[k].squeeze().cpu().numpy()
        job.results = _score(job.args, p, np.asarray(job.labels[2]))
        job.train_losses = [float(v) for v in train_sums[:, k]]
        job.valid_losses = [float(v) for v in valid_vals[:, k]]
    return jobs


def run_jobs(jobs, device, max_batch=16):
    """Group the jobs by shape and train each group batched; returns them in the order given."""
    groups = {}
    for j in jobs:
        groups.setdefault(j.key(), []).append(j)
    for group in groups.values():
        for o in range(0, len(group), max_batch):
            train_batched(group[o:o + max_batch], device)
    return jobs
