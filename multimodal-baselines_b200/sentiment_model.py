"""Downstream sentiment regressor (reference sentiment_model.py).

Outside the hot path by the task's own scope ("unchanged, only used for parity checks"): a
small PyTorch MLP trained with L1 loss on fixed latents.  Same class / function names as the
reference so that ``simplesif`` keeps its imports; no custom kernels here.
"""
import json
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim as optim
from torch.utils.data import DataLoader, Dataset

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


def predict_sentiment(data, model, latents):
    """reference sentiment_model.py:51-74 -- predictions and targets as NumPy arrays."""
    ys, ps = [], []
    with torch.no_grad():
        for j, senti in data:
            ys.append(senti)
            ps.append(model(latents[j]).reshape(senti.shape))
    y, p = torch.cat(ys), torch.cat(ps)
    print("MAE: {}".format(float((p - y).abs().sum() / len(data.dataset))))
    return p.cpu().numpy(), y.cpu().numpy()


class _GraphedSentimentStep(object):
    """One SGD step of the regressor (forward, L1, backward, update) captured as a CUDA graph per
    batch size and replayed; numerically the eager step on the same indices."""

    def __init__(self, model, latents, labels, optimizer):
        self.model, self.latents, self.labels, self.optimizer = model, latents, labels, optimizer
        self.graphs = {}

    def _step(self, j):
        senti = self.labels[j]
        loss = (self.model(self.latents[j]).reshape(senti.shape) - senti).abs().mean()
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
            with torch.cuda.graph(graph):
                static_loss = self._step(static_j)
            self.graphs[n] = (graph, static_j, static_loss)
        graph, static_j, static_loss = self.graphs[n]
        static_j.copy_(j)
        graph.replay()
        return static_loss


def train_sentiment(args, model, train_data, train_latents, valid_data=None, valid_latents=None,
                    model_save_path=None):
    """reference sentiment_model.py:76-163 -- SGD on the L1 loss; with ``early_stopping`` the best
    validation checkpoint is restored and the step size decayed by ``lr_decay`` on plateaus."""
    lr = args['sentiment_lr']
    optimizer = optim.SGD(model.parameters(), lr=lr)
    best, best_state, patience, trials = float('inf'), None, 0, 0
    train_losses, valid_losses = [], []
    graphed = str(args.get('cuda_graph', 0)) not in ('0', 'False', 'false', '') and train_latents.is_cuda
    stepper = _GraphedSentimentStep(model, train_latents, train_data.dataset.sentiment, optimizer) if graphed else None
    for _ in range(args['n_sentiment_epochs']):
        if graphed:
            # each SGD step is one graph replay; the epoch loss is read back once
            total_t = torch.zeros((), device=train_latents.device)
            torch.empty((), dtype=torch.int64).random_(generator=train_data.generator)   # the DataLoader's base seed
            for idx in train_data.batch_sampler:
                total_t += stepper(torch.as_tensor(idx, dtype=torch.int64).to(train_latents.device, non_blocking=True))
            total = float(total_t)
        else:
            total = 0.
            for j, senti in train_data:
                optimizer.zero_grad()
                loss = (model(train_latents[j]).reshape(senti.shape) - senti).abs().mean()
                loss.backward()
                optimizer.step()
                total += float(loss)
        train_losses.append(total)
        if args.get('early_stopping') and valid_data is not None:
            with torch.no_grad():
                v = sum(float((model(valid_latents[j]).reshape(s.shape) - s).abs().sum()) for j, s in valid_data)
            valid_losses.append(v)
            if v < best:
                best, patience = v, 0
                best_state = {k: t.clone() for k, t in model.state_dict().items()}
            else:
                patience += 1
                if patience >= 5:
                    trials, patience = trials + 1, 0
                    if trials >= 5:
                        print("early stopping...")
                        break
                    lr = lr * args.get('lr_decay', 0.5)
                    model.load_state_dict(best_state)
                    optimizer = optim.SGD(model.parameters(), lr=lr)
                    if graphed:
                        stepper = _GraphedSentimentStep(model, train_latents, train_data.dataset.sentiment, optimizer)
    if best_state is not None:
        model.load_state_dict(best_state)
    return train_losses, valid_losses


def _score(args, predictions, y):
    if args['dataset'] == 'mosi':
        return full_loss(predictions, y)
    if args['dataset'] == 'pom':
        return pom_loss(predictions, y)
    return iemocap_loss(predictions, y)


def train_sentiment_for_latents(args, latents, sentiment_data, device, train_idxes=None, model_save_path=None):
    """reference sentiment_model.py:165-265 -- fit the regressor on the train latents, report the
    dataset's metrics on the test latents, write ``test_results_after.json``."""
    train_latents, valid_latents, test_latents = latents
    train_s, valid_s, test_s = sentiment_data
    n_out = 1 if train_s.ndim == 1 else train_s.shape[-1]
    if train_idxes is not None:
        train_latents, train_s = train_latents[train_idxes], train_s[train_idxes]
    loaders = [DataLoader(SentimentData(s, device), batch_size=32, shuffle=sh)
               for s, sh in ((train_s, True), (valid_s, False), (test_s, False))]
    model = SentimentModel(train_latents.shape[-1], args['sentiment_hidden_size'], n_out).to(device)
    losses_ = train_sentiment(args, model, loaders[0], train_latents, loaders[1], valid_latents, model_save_path)
    predictions, y = predict_sentiment(loaders[2], model, test_latents)
    results = _score(args, predictions, y)
    if model_save_path is not None:
        with open(os.path.join(model_save_path, 'test_results_after.json'), 'w') as f:
            json.dump(results, f, indent=2)
    return results, losses_
