"""Downstream sentiment regressor (reference sentiment_model.py).

Outside the hot path by the task's own scope ("unchanged, only used for parity checks"): a small
PyTorch MLP trained with L1 loss on fixed latents, no custom kernels.  Same names, argument order,
training schedule, files written and -- because the downstream metrics are compared with the
reference's to the third decimal (tests/test_mmb_gpu.py::test_downstream_metrics) -- the same
consumption of torch's global random stream: three shuffled loaders of batch 32, an initial pass
over the test loader, a pass over the validation loader every ``valid_niter`` epochs whether or not
early stopping is on.

Two things differ in mechanism, not in result: batches are taken as index vectors from the
DataLoader's own batch sampler (same draws, no per-sample collation of device scalars), and with
``args['cuda_graph']`` each SGD step is one CUDA-graph replay.
"""
import json
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim
from torch.utils.data import DataLoader, Dataset

import graph_capture
from losses import full_loss, iemocap_loss, pom_loss


class SentimentData(Dataset):
    """reference sentiment_model.py:14-27 -- ``__getitem__`` returns ``(idx, label[idx])``."""

    def __init__(self, sentiment, device):
        super(Dataset, self).__init__()
        if not torch.is_tensor(sentiment):
            sentiment = torch.tensor(sentiment, device=device, dtype=torch.float32)
        self.sentiment = sentiment

    def __len__(self):
        return self.sentiment.size()[0]

    def __getitem__(self, idx):
        return idx, self.sentiment[idx]


class SentimentModel(nn.Module):
    """reference sentiment_model.py:29-41 -- Linear -> ReLU -> Linear, squeezed output."""

    def __init__(self, embedding_dim, hidden_dim, n_out):
        super(SentimentModel, self).__init__()
        self.hidden1 = nn.Linear(embedding_dim, hidden_dim)
        self.out = nn.Linear(hidden_dim, n_out)

    def forward(self, inputs):
        return self.out(F.relu(self.hidden1(inputs))).squeeze()


def save_sentiment(path, model):
    """reference sentiment_model.py:43-44."""
    torch.save(model.state_dict(), os.path.join(path, 'senti.bin'))


def load_sentiment(path, embedding_dim, hidden_dim, device, n_out=1):
    """reference sentiment_model.py:46-50 (which omits ``n_out`` and cannot run; only ever bound in an
    unused lambda, line 226)."""
    model = SentimentModel(embedding_dim, hidden_dim, n_out)
    model.load_state_dict(torch.load(path))
    return model.to(device)


def _index_batches(loader, device):
    """The index batches ``for j, senti in loader`` would yield, with the same draws from the
    global generator in the same order (num_workers = 0: the iterator's base seed first, then the
    RandomSampler's seed when the first batch is requested)."""
    torch.empty((), dtype=torch.int64).random_(generator=loader.generator)
    batches = list(loader.batch_sampler)
    if not batches:
        return
    # one host-to-device copy per pass, not one per batch
    flat = torch.tensor([i for b in batches for i in b], dtype=torch.int64).to(device, non_blocking=True)
    off = 0
    for b in batches:
        yield flat[off:off + len(b)]
        off += len(b)


def _epoch_indices(loader, device):
    """``_index_batches`` as one flat index tensor + the batch sizes (same draws from the generator)."""
    torch.empty((), dtype=torch.int64).random_(generator=loader.generator)
    batches = list(loader.batch_sampler)
    flat = torch.tensor([i for b in batches for i in b], dtype=torch.int64).to(device, non_blocking=True)
    return flat, [len(b) for b in batches]


def _l1(model, latents, labels, j):
    # nn.L1Loss(reduce=False) of the reference (sentiment_model.py:90, 103): plain broadcasting of the
    # squeezed prediction against the label batch.  For (B,) or (B, n_out) labels that is element-wise;
    # for (B, 1) labels the reference's (B,) prediction broadcasts to (B, B) -- reproduced, not "fixed",
    # because the downstream MAE / correlation of such a run depend on it.
    senti = labels[j]
    return (model(latents[j]) - senti).abs(), senti


def predict_sentiment(data, model, latents):
    """reference sentiment_model.py:51-74 -- predictions and targets as NumPy arrays (in the
    loader's shuffled order, like the reference; the metrics do not depend on the order)."""
    labels = data.dataset.sentiment
    ys, ps = [], []
    total = torch.zeros((), device=labels.device)
    with torch.no_grad():
        for j in _index_batches(data, labels.device):
            senti = labels[j]
            p = model(latents[j])                    # squeezed, as the reference feeds it to L1Loss and cat
            total += (p - senti).abs().sum()         # broadcasts like nn.L1Loss(reduce=False)
            ys.append(senti)
            ps.append(p.reshape(1) if p.dim() == 0 else p)   # (a batch of one: the reference's cat would raise)
    print("MAE: {}".format(float(total) / len(data.dataset)))
    return torch.cat(ps).cpu().numpy(), torch.cat(ys).cpu().numpy()


class _GraphedSentimentStep(object):
    """One SGD step of the regressor (forward, L1, backward, update) captured as a CUDA graph per
    batch size and replayed; numerically the eager step on the same indices."""

    def __init__(self, model, latents, labels, optimizer):
        self.model, self.latents, self.labels, self.optimizer = model, latents, labels, optimizer
        self.graphs = {}
        self.epoch_graphs = {}

    def _step(self, j):
        loss = _l1(self.model, self.latents, self.labels, j)[0].mean()
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, j):
        n = int(j.shape[0])
        if n not in self.graphs:
            dev = self.latents.device
            static_j = torch.zeros(n, dtype=torch.int64, device=dev)
            params = [p for g in self.optimizer.param_groups for p in g['params']]
            saved = [p.detach().clone() for p in params]
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self.optimizer.zero_grad(set_to_none=True)
                    self._step(static_j)
            torch.cuda.current_stream(dev).wait_stream(side)
            with torch.no_grad():
                for p, s in zip(params, saved):
                    p.copy_(s)
            graph = torch.cuda.CUDAGraph()
            self.optimizer.zero_grad(set_to_none=True)
            with graph_capture.capture(graph):
                static_loss = self._step(static_j)
            self.graphs[n] = (graph, static_j, static_loss)
        graph, static_j, static_loss = self.graphs[n]
        static_j.copy_(j)
        graph.replay()
        return static_loss

    def run_epoch(self, flat, sizes):
        """All SGD steps of one epoch as ONE graph replay (the batches are static slices of one index
        buffer); returns the (device) sum of the per-step losses."""
        key = tuple(int(n) for n in sizes)
        if key not in self.epoch_graphs:
            dev = self.latents.device
            static_flat = torch.zeros(int(sum(key)), dtype=torch.int64, device=dev)
            params = [p for g in self.optimizer.param_groups for p in g['params']]
            saved = [p.detach().clone() for p in params]
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for n in sorted(set(key)) * 2:
                    self.optimizer.zero_grad(set_to_none=True)
                    self._step(static_flat[:n])
            torch.cuda.current_stream(dev).wait_stream(side)
            with torch.no_grad():
                for p, s_ in zip(params, saved):
                    p.copy_(s_)
            graph = torch.cuda.CUDAGraph()
            self.optimizer.zero_grad(set_to_none=True)
            with graph_capture.capture(graph):
                total = torch.zeros((), device=dev)
                off = 0
                for n in key:
                    self.optimizer.zero_grad(set_to_none=True)
                    total = total + self._step(static_flat[off:off + n])
                    off += n
            self.epoch_graphs[key] = (graph, static_flat, total)
        graph, static_flat, total = self.epoch_graphs[key]
        static_flat.copy_(flat)
        graph.replay()
        return total


def train_sentiment(args, model, train_data, train_latents, valid_data, valid_latents, model_loader=None,
                    valid_niter=10, verbose=False, model_save_path=None):
    """reference sentiment_model.py:76-163 -- plain SGD on the mean L1 loss, batches of the shuffled
    loader; every ``valid_niter`` epochs one pass over the validation loader.  With
    ``args['early_stopping']``: patience 10 validations, then reload the best checkpoint (model +
    optimizer, when a save path is given) and multiply the step size by ``args['lr_decay']``, at most
    3 times.  Returns ``(train_losses, valid_losses)`` (per-epoch / per-validation batch means)."""
    n_epochs = args['n_sentiment_epochs']
    lr = args['sentiment_lr']
    patience, n_trials = 10, 3
    n_samples = len(train_data.dataset)
    device = train_latents.device
    labels, valid_labels = train_data.dataset.sentiment, valid_data.dataset.sentiment
    optimizer = optim.SGD(model.parameters(), lr=lr)
    graphed = str(args.get('cuda_graph', 0)) not in ('0', 'False', 'false', '') and train_latents.is_cuda
    stepper = _GraphedSentimentStep(model, train_latents, labels, optimizer) if graphed else None
    graph_epochs = graphed and str(args.get('cuda_graph')) != 'step'    # one graph per epoch, else one per step
    ckpt_file = os.path.join(model_save_path, 'senti.bin') if model_save_path is not None else None

    train_losses, valid_losses = [], []
    n_bad = n_bad_trials = 0
    i, epoch_loss = -1, torch.zeros((), device=device)
    for i in range(n_epochs):
        epoch_loss = torch.zeros((), device=device)
        n_batches = 0
        if graph_epochs:
            flat, sizes = _epoch_indices(train_data, device)
            n_batches = len(sizes)
            if n_batches:
                epoch_loss = stepper.run_epoch(flat, sizes)
        else:
            for j in _index_batches(train_data, device):
                n_batches += 1
                if graphed:
                    epoch_loss += stepper(j)
                else:
                    model.zero_grad()
                    loss = _l1(model, train_latents, labels, j)[0].mean()
                    loss.backward()
                    optimizer.step()
                    epoch_loss += loss.detach()
        # graph-captured epochs: keep the epoch's sum on the device (a clone: the next replay overwrites the
        # graph's static output) and read all of them back once at the end -- no per-epoch stall
        train_losses.append((epoch_loss.clone() if graph_epochs else float(epoch_loss), max(n_batches, 1)))
        if i % valid_niter == 0:
            batches = 0
            valid_loss = torch.zeros((), device=device)
            with torch.no_grad():
                for j in _index_batches(valid_data, device):
                    valid_loss += _l1(model, valid_latents, valid_labels, j)[0].mean()
                    batches += 1
            avg_valid_loss = float(valid_loss) / max(batches, 1)
            print("Epoch {}: {} (avg val loss {})".format(i, float(train_losses[-1][0]) / train_losses[-1][1],
                                                          avg_valid_loss))
            is_better = len(valid_losses) == 0 or avg_valid_loss < min(valid_losses)
            valid_losses.append(avg_valid_loss)
            if args['early_stopping']:
                if is_better:
                    n_bad = 0
                    if ckpt_file is not None:
                        torch.save({'model_state_dict': model.state_dict(),
                                    'optimizer_state_dict': optimizer.state_dict()}, ckpt_file)
                else:
                    print('patience {}'.format(n_bad))
                    n_bad += 1
                    if n_bad >= patience:
                        n_bad_trials += 1
                        if n_bad_trials < n_trials:
                            if ckpt_file is not None:
                                print("reloading model and decaying learning rate...")
                                checkpoint = torch.load(ckpt_file)
                                model.load_state_dict(checkpoint['model_state_dict'])
                                optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
                            lr = lr * args['lr_decay']
                            for g in optimizer.param_groups:
                                g['lr'] = lr
                            if graphed:   # the step size is baked into the captured update
                                stepper = _GraphedSentimentStep(model, train_latents, labels, optimizer)
                            n_bad = 0
                        else:
                            print("early stopping...")
                            break
    print("Epoch {}: {}".format(i, float(epoch_loss) / max(n_samples, 1)))
    if train_losses and any(torch.is_tensor(v) for v, _ in train_losses):
        sums = torch.stack([v if torch.is_tensor(v) else torch.tensor(v, device=device) for v, _ in train_losses]).cpu()
        train_losses = [float(v) / n for v, (_, n) in zip(sums, train_losses)]
    else:
        train_losses = [float(v) / n for v, n in train_losses]
    return train_losses, valid_losses


def _score(args, predictions, y):
    if args['dataset'] == 'mosi':
        return full_loss(predictions, y)
    elif args['dataset'] == 'iemocap':
        return iemocap_loss(predictions, y)
    return pom_loss(predictions, y)


def train_sentiment_for_latents(args, latents, sentiment_data, device, verbose=False, model_save_path=None,
                                train_idxes=None):
    """reference sentiment_model.py:165-265 -- fit the regressor on the train latents, score the test
    latents before and after with the dataset's metrics, write ``test_results_{before,after}.json``,
    ``test_acc_*.txt``, ``senti_{train,valid}_loss.txt`` and ``senti.bin`` under ``model_save_path``.
    Like the reference, the final scores come from the model as training left it (the reference
    loads its best checkpoint into a second model that it never uses, lines 247-251).
    The reference returns None; this returns ``(results_after, (train_losses, valid_losses))``."""
    train_latents, valid_latents, test_latents = latents
    hidden_dim = args['sentiment_hidden_size']
    embedding_dim = train_latents.size()[-1]
    train, valid, test = sentiment_data
    n_out = 1 if train.ndim == 1 else train.shape[-1]
    senti_model = SentimentModel(embedding_dim, hidden_dim, n_out).to(device)

    print("train data shape:", train.shape)
    print("train latents shape:", train_latents.size())
    if train_idxes is not None:
        train = train[train_idxes]
        train_latents = train_latents[train_idxes]
        print("train data shape:", train.shape)
        print("train latents shape:", train_latents.size())

    train_data, valid_data, test_data = (SentimentData(s, device) for s in (train, valid, test))
    assert train_latents.size()[0] == train.shape[0]
    print("# of sentiment points:", len(train_data))
    train_loader = DataLoader(train_data, batch_size=32, shuffle=True)
    valid_loader = DataLoader(valid_data, batch_size=32, shuffle=True)
    test_loader = DataLoader(test_data, batch_size=32, shuffle=True)

    print("Initial sentiment predictions")
    senti_model.eval()
    predictions, y_test = predict_sentiment(test_loader, senti_model, test_latents)
    results = _score(args, predictions, y_test)
    if model_save_path is not None:
        if 'accuracy' in results:
            with open(os.path.join(model_save_path, 'test_acc_before.txt'), 'w') as f:
                f.write(str(results['accuracy']))
        with open(os.path.join(model_save_path, 'test_results_before.json'), 'w') as f:
            json.dump(results, f, indent=2)

    print("Training sentiment model on sentence embeddings...")
    senti_model.train()
    train_losses, valid_losses = train_sentiment(args, senti_model, train_loader, train_latents, valid_loader,
                                                 valid_latents, None, verbose=verbose,
                                                 model_save_path=model_save_path)
    if model_save_path is not None:
        for name, vals in (('senti_train_loss.txt', train_losses), ('senti_valid_loss.txt', valid_losses)):
            with open(os.path.join(model_save_path, name), 'w') as f:
                f.writelines('{}\n'.format(v) for v in vals)
        if not args['early_stopping']:
            save_sentiment(model_save_path, senti_model)
    if args['early_stopping']:
        # the reference builds a second model here for the best checkpoint and then scores the first
        # one (lines 247-251); only its draws from the global generator have any effect
        print('reloading best')
        SentimentModel(embedding_dim, hidden_dim, n_out)

    print("Sentiment predictions after training")
    senti_model.eval()
    predictions, y_test = predict_sentiment(test_loader, senti_model, test_latents)
    results = _score(args, predictions, y_test)
    if model_save_path is not None:
        if 'accuracy' in results:
            with open(os.path.join(model_save_path, 'test_acc_after.txt'), 'w') as f:
                f.write(str(results['accuracy']))
        with open(os.path.join(model_save_path, 'test_results_after.json'), 'w') as f:
            json.dump(results, f, indent=2)
    print("-----------------------------")
    return results, (train_losses, valid_losses)
