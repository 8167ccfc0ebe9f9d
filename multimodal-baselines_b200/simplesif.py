"""Drop-in for the hot-path part of the reference's ``simplesif.py``.

Kept with the reference's names and signatures: ``update_masks``, ``update_masks_vect``,
``optimize_latents`` (reference simplesif.py:36-162), ``read_config`` / ``parse_arguments``
(177-238) and ``main`` (240-916).  The SIF initialisation goes through ``sif.py`` and the
latent-optimisation loop through ``models.py`` / ``losses.py``, i.e. through libmmb_b200.so.

What changes inside the loop (results do not):
  * the four ``torch.cat`` of data and four of masks per step (reference 94-113) are
    ``losses.CatSegments`` -- the fused Gaussian kernel reads the base tensors directly;
  * the per-modality ``sigma.min()`` / ``lp.min()`` device syncs (reference 82-84 and
    losses.py:258-264, ~14 per MMB2 step) collapse into one status word per step.
The reference imports ``analyze_embeddings.get_closest_words``, a module that is not in its
tree; it is optional here.
"""
import argparse
import json
import os
import pprint
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim
from torch.utils.data import DataLoader

from losses import CatSegments, MomentStats, TokenIds, get_log_prob_matrix, get_word_log_prob_angular, get_word_log_prob_dot_prod  # noqa: F401
from losses import get_word_log_prob_angular2
import graph_capture
import mmb_ops
from models import AudioVisualGeneratorConcat, AudioVisualGenerator, AudioVisualGeneratorMultimodal  # noqa: F401
from sentiment_model import SentimentData, SentimentModel, train_sentiment_for_latents
from sif import load_weights, get_sentence_embeddings
from utils import load_data, normalize_data, MMData, MMDataExtra, MMDataIds, MMDataExtraIds, add_positional_embeddings

try:  # not part of the reference tree either
    from analyze_embeddings import get_closest_words
except ImportError:  # pragma: no cover
    get_closest_words = None


def update_masks(mask_dict, data, embedding_dim):
    """reference simplesif.py:36-40 -- text mask = (ids != 0) broadcast to (N, L, embedding_dim)."""
    tmp = (data != 0).astype(int)
    mask_dict['text'] = np.broadcast_to(np.expand_dims(tmp, -1), tmp.shape + (embedding_dim,))
    print(np.all(mask_dict['text'][:, :, 1] == mask_dict['text'][:, :, 0]))


def update_masks_vect(mask_dict, data, key='text'):
    """reference simplesif.py:42-47 -- a time step is valid iff no feature of it is exactly 0."""
    tmp2 = np.all(data != 0, axis=-1).astype(int)
    print(tmp2.shape)
    mask_dict[key] = np.broadcast_to(np.expand_dims(tmp2, -1), data.shape)


def update_masks_device(mask_dict, ids, embedding_dim, device=None):
    """``update_masks`` (reference simplesif.py:36-40) on the device, SURVEY.md 8f N3: ``mask_dict['text']``
    becomes the (N, L, embedding_dim) float32 CUDA mask as a stride-0 VIEW of the (N, L) vector ``ids != 0``
    (``mmb_token_mask``) -- the reference materialises N * L * embedding_dim ints on the host (1 GB at POM).
    ``MMData`` and the kernels take the view as is (only ``mask[:, :, 0]`` is ever read for the text)."""
    import _native as nv
    dev = device or nv.require_cuda()
    ids_t = nv.to_device(ids, torch.int64, dev)
    m = torch.empty(ids_t.shape, dtype=torch.float32, device=dev)
    nv.check(nv.lib.mmb_token_mask(nv.ptr(ids_t), ids_t.numel(), nv.ptr(m), nv.stream_ptr()))
    mask_dict['text'] = m.unsqueeze(-1).expand(*ids_t.shape, embedding_dim)
    return mask_dict['text']


def update_masks_vect_device(mask_dict, data, key='text', device=None):
    """``update_masks_vect`` (reference simplesif.py:42-47) on the device: a time step is valid iff no
    feature of it is exactly 0 (``mmb_step_mask``); the (N, T, F) mask is a stride-0 view of the (N, T) one."""
    import _native as nv
    dev = device or nv.require_cuda()
    x = nv.to_device(data, torch.float32, dev)
    N, T, F = x.shape
    m = torch.empty((N, T), dtype=torch.float32, device=dev)
    nv.check(nv.lib.mmb_step_mask(nv.ptr(x), N * T, F, nv.ptr(m), nv.stream_ptr()))
    mask_dict[key] = m.unsqueeze(-1).expand(N, T, F)
    return mask_dict[key]


def _batch_dicts(args, x, table=None):
    """The per-step ``batch_data`` / ``batch_masks`` of reference simplesif.py:72-124, with the
    concatenated modalities expressed as CatSegments instead of materialised torch.cat.  A batch of an
    id-based dataset (``MMDataIds``: int64 ids in the text slot) is wrapped as ``TokenIds(ids, table)``."""
    text_moments = None
    if args['dataset'] == 'mosi':
        if len(x) == 9:                    # _with_moments: the text's moments ride behind the tuple
            x, text_moments = x[:8], x[8]
        j, text, aud, vis, text_m, aud_m, vis_m, text_w = x
    else:
        j, text, aud, vis, text_m, aud_m, vis_m, text_w, text_gauss, text_gauss_m = x
    if torch.is_tensor(text) and text.dtype == torch.int64:
        text = TokenIds(text, table)
    if args['dataset'] == 'mosi':
        text_gauss, text_gauss_m = (text, text_m) if text_moments is None else (text_moments, None)
    batch_data = {'text': text, 'audio': aud, 'visual': vis, 'text_weights': text_w}
    batch_masks = {'text': text_m, 'audio': aud_m, 'visual': vis_m}
    if not args['unimodal']:
        batch_data.update({
            'audiovisual': CatSegments([aud, vis]),
            'textaudio': CatSegments([text_gauss, aud]),
            'textvisual': CatSegments([text_gauss, vis]),
            'textaudiovisual': CatSegments([text_gauss, aud, vis]),
        })
        batch_masks.update({
            'audiovisual': CatSegments([aud_m, vis_m]),
            'textaudio': CatSegments([text_gauss_m, aud_m]),
            'textvisual': CatSegments([text_gauss_m, vis_m]),
            'textaudiovisual': CatSegments([text_gauss_m, aud_m, vis_m]),
        })
    return j, batch_data, batch_masks


def _dataset_moments(args, dataset):
    """The per-utterance moments of every tensor the Gaussian terms read, computed
    once per dataset and cached on it: the data of an utterance is the same at every step of every epoch, so
    the loops hand ``get_log_prob_matrix`` 3 numbers per (utterance, feature) instead of the (B, T, F) values
    and masks.  ``args['gauss_moments'] = 0`` (or MMB_GAUSS_MOMENTS=0) keeps the raw tensors."""
    flag = args.get('gauss_moments', os.environ.get('MMB_GAUSS_MOMENTS', '1'))
    if str(flag) in ('0', 'False', 'false', '') or not hasattr(dataset, 'audio_mask'):
        return None
    if not dataset.audio.is_cuda:
        return None
    cached = getattr(dataset, '_gauss_moments', None)
    key = (args['dataset'], bool(args['unimodal']))
    if cached is not None and cached[0] == key:
        return cached[1]
    mom = {'audio': MomentStats.of(dataset.audio, dataset.audio_mask),
           'visual': MomentStats.of(dataset.visual, dataset.visual_mask)}
    if not args['unimodal']:
        if args['dataset'] != 'mosi':
            mom['text_gauss'] = MomentStats.of(dataset.text_aligned, dataset.text_aligned_mask)
        elif hasattr(dataset, 'text_ids'):
            mom['text_gauss'] = MomentStats.of(TokenIds(dataset.text_ids, dataset.table), dataset.text_mask)
        else:
            mom['text_gauss'] = MomentStats.of(dataset.text, dataset.text_mask)
    dataset._gauss_moments = (key, mom)
    return mom


def _with_moments(args, x, moments):
    """A batch tuple (reference utils.py:231-233 / 248-251) with the Gaussian inputs replaced by the moments
    of its rows; the word term's inputs (text, text mask, weights) are untouched."""
    if moments is None:
        return x
    x = list(x)
    j = x[0]
    x[2], x[3], x[5], x[6] = moments['audio'][j], moments['visual'][j], None, None
    if 'text_gauss' in moments:
        if args['dataset'] == 'mosi':
            x.append(moments['text_gauss'][j])     # consumed by _batch_dicts as the Gaussian view of the text
        else:
            x[8], x[9] = moments['text_gauss'][j], None
    return tuple(x)


def _make_optimizer(args, params, lr):
    if args['optimizer'] == 'sgd':
        return optim.SGD(params, lr=lr)
    elif args['optimizer'] == 'adam':
        return optim.Adam(params, lr=lr)
    raise NotImplementedError(args['optimizer'])


def _warn_tiny_sigma(out):
    """reference simplesif.py:82-84, one device sync for all modalities instead of one each."""
    smallest = torch.stack([d['sigma'].detach().min() for d in out.values()]).min()
    if float(smallest.abs()) < 1e-7:
        print({m: float(d['sigma'].min()) for m, d in out.items()}, "boo!")


class _RowGather(torch.autograd.Function):
    """``embeddings[j]`` with a leaner backward for the captured step: the dense gradient the optimizers need
    (reference simplesif.py:54-61 steps the whole (N, d) latent tensor) is a zero fill plus ``index_add_``
    -- two launches instead of the sort-based ``index_put_(accumulate=True)`` chain of advanced indexing
    (six).  Same values: a DataLoader batch never repeats a row, and repeated rows would still be summed."""

    @staticmethod
    def forward(ctx, table, j):
        ctx.save_for_backward(j)
        ctx.n = table.shape[0]
        return table.index_select(0, j)

    @staticmethod
    def backward(ctx, g):
        (j,) = ctx.saved_tensors
        grad = torch.zeros((ctx.n,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        grad.index_add_(0, j, g)
        return grad, None


CAPTURE_SECONDS = [0.0]      # wall time spent warming up + capturing graphs (sweep.py reports it per config)


class GraphedStep(object):
    """One latent-optimisation step -- gather the batch, generator heads, word + Gaussian
    log-likelihoods, backward, optimizer step -- captured ONCE as a CUDA graph and replayed per
    batch.  The step is a few dozen small launches at B = 64, so in the
    eager loop Python, launch latency and the per-step status read-back dominate; the replay is
    one launch with no host synchronisation (non-finite values are flagged in a device word that
    the caller checks once per epoch).

    The captured work is exactly what the eager loop of ``optimize_latents`` runs on the same
    indices, so the results agree to rounding (tests/test_mmb_gpu.py)."""

    def __init__(self, args, gen_model, embeddings, dataset, optimizer, word_prob_fn, device, extra_loss=None,
                 extra_modules=()):
        """``extra_loss(j, latents_j, neg_log_prob)`` -> per-utterance loss (the e2e loop mixes in the
        sentiment regressor's L1 term); ``extra_modules`` are restored with the rest after warm-up."""
        import mmb_ops
        self.extra_loss, self.extra_modules = extra_loss, list(extra_modules)
        self.args, self.gen_model, self.embeddings = args, gen_model, embeddings
        self.dataset, self.optimizer, self.word_prob_fn, self.device = dataset, optimizer, word_prob_fn, device
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.graphs = {}
        self.epoch_graphs = {}
        self._keepalive = []
        self.mmb_ops = mmb_ops
        self.moments = _dataset_moments(args, dataset)
        # single-step graphs fork the word term (0.149 -> 0.124 ms per MOSI step); whole-epoch graphs do not by
        # default: with 21 forked steps chained per graph and the graphs re-captured for every grid point, the
        # 4-config sweep took 16.4-17.0 s forked against 12.5-14.9 s unforked on the same box
        off = ('0', 'False', 'false', '')
        self.fork = str(args.get('graph_fork', os.environ.get('MMB_GRAPH_FORK', '1'))) not in off
        self.fork_epoch = str(args.get('graph_fork_epoch', os.environ.get('MMB_GRAPH_FORK_EPOCH', '0'))) not in off
        self.side = torch.cuda.Stream(device=device) if (self.fork or self.fork_epoch) else None
        # Step-cache mode (args['_step_cache'], sweep.py): this stepper outlives the call and serves later grid points
        # of the same structure.  Whatever differs between them must then live in device memory, not in the captured
        # launches: the likelihood weights of losses.py:267-272 (and, for the e2e loop, of simplesif.py:786) become
        # one-element tensors that ``rebind`` rewrites.  Same float32 values, same kernels, same rounding.
        self.cache_mode = args.get('_step_cache') is not None
        if self.cache_mode:
            self.args = dict(args)
            self.loss_w = torch.zeros(4, dtype=torch.float32, device=device)    # other_w, word_w, like_w, 1 - like_w
            self.args['_loss_weights_dev'] = (self.loss_w[0:1], self.loss_w[1:2]) if 'word_loss_weight' in args else None
            self._set_weights(args)

    def _set_weights(self, args):
        n_mod = len(self.gen_model.embed2out)
        ww = float(args['word_loss_weight']) if 'word_loss_weight' in args else 1.
        ow = (1. - ww) / n_mod if 'word_loss_weight' in args else 1.
        lw = float(args.get('likelihood_weight', 1.))
        self.loss_w.copy_(torch.tensor([ow, ww, lw, 1. - lw], dtype=torch.float64).to(torch.float32), non_blocking=True)

    def rebind(self, args, embed_init):
        """Serve another call with the captured graphs: new initial latents, new likelihood weights, optimizer
        state back to zero (the generator / regressor parameters were re-initialised in place by the caller)."""
        with torch.no_grad():
            self.embeddings.copy_(torch.as_tensor(np.asarray(embed_init), dtype=torch.float32), non_blocking=True)
            for st in self.optimizer.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        self._set_weights(args)
        self.status.zero_()

    def _gather(self, j):
        """``dataset[j]`` (reference utils.py:231-233 / 248-251) as one multi-tensor gather launch."""
        ds = self.dataset
        mom = self.moments
        if mom is not None:
            # word-term inputs + the moments of the Gaussian inputs: one multi-tensor launch of small rows
            keys = ['audio', 'visual'] + (['text_gauss'] if 'text_gauss' in mom else [])
            srcs = ([] if hasattr(ds, 'text_ids') else [ds.text]) + [ds.text_mask, ds.text_weights] + \
                [mom[k].stats for k in keys]
            got = list(self.mmb_ops.gather_multi(srcs, j))
            text = ds.text_ids[j] if hasattr(ds, 'text_ids') else got.pop(0)
            text_m, text_w = got[0], got[1]
            st = {k: MomentStats(t) for k, t in zip(keys, got[2:])}
            x = (j, text, st['audio'], st['visual'], text_m, None, None, text_w)
            if 'text_gauss' in st:
                x = x + ((st['text_gauss'],) if self.args['dataset'] == 'mosi' else (st['text_gauss'], None))
            elif self.args['dataset'] != 'mosi':
                x = x + (None, None)
            return x
        names = ['text', 'audio', 'visual', 'text_mask', 'audio_mask', 'visual_mask', 'text_weights']
        if hasattr(ds, 'text_aligned'):
            names += ['text_aligned', 'text_aligned_mask']
        if hasattr(ds, 'text_ids'):      # id-based dataset: the text slot is an int64 gather of (B, L) ids
            got = self.mmb_ops.gather_multi([getattr(ds, n) for n in names[1:]], j)
            return (j, ds.text_ids[j]) + tuple(got)
        return (j,) + tuple(self.mmb_ops.gather_multi([getattr(ds, n) for n in names], j))

    def _step(self, j, fork=None):
        fork = self.fork if fork is None else fork
        x = self._gather(j)                       # batched gather of the device-resident tensors
        _, batch_data, batch_masks = _batch_dicts(self.args, x, getattr(self.dataset, 'table', None))
        e = _RowGather.apply(self.embeddings, j)
        # The word term (vocabulary-sized products) and the heads + Gaussian terms only meet in the final sum,
        # and each of their kernels fills a fraction of the GPU.  The word term is issued first -- with `fork`
        # on a second stream, forked from and joined back into this one, so that the captured graph has two
        # parallel branches (autograd runs each branch's backward on the stream of its forward).  The order
        # of the autograd nodes, hence every rounding, is the same with and without the fork.
        main = torch.cuda.current_stream(self.device)
        if fork:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                word_lp = self.word_prob_fn(e, batch_data['text_weights'], batch_data['text'], batch_masks['text'])
        else:
            word_lp = self.word_prob_fn(e, batch_data['text_weights'], batch_data['text'], batch_masks['text'])

        def word_fn(latents, word_weights, sent_embeddings, mask):
            if fork:
                main.wait_stream(self.side)
                word_lp.record_stream(main)
            return word_lp
        out = self.gen_model(e)
        log_prob = -get_log_prob_matrix(self.args, e, out, batch_data, batch_masks, word_fn,
                                        device=self.device, verbose=False)
        if self.extra_loss is not None:
            log_prob = self.extra_loss(j, e, log_prob)
        loss = log_prob.mean()
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _snapshot(self):
        params = [p for g in self.optimizer.param_groups for p in g['params']]
        saved = [p.detach().clone() for p in params]
        bufs = [b.detach().clone() for b in self._buffers()]
        return params, saved, bufs

    def _buffers(self):
        out = list(self.gen_model.buffers())
        for m in self.extra_modules:
            out.extend(m.buffers())
        return out

    def _restore(self, params, saved, bufs):
        with torch.no_grad():
            for p, s in zip(params, saved):
                p.copy_(s)
            for b, s in zip(self._buffers(), bufs):
                b.copy_(s)
            for st in self.optimizer.state.values():     # Adam moments / step counters back to zero
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()

    def _capture(self, n):
        """Warm up on a side stream (state restored afterwards), then capture a step for n rows."""
        t0 = time.perf_counter()
        try:
            return self._capture_impl(n)
        finally:
            CAPTURE_SECONDS[0] += time.perf_counter() - t0

    def _capture_impl(self, n):
        static_j = torch.zeros(n, dtype=torch.int64, device=self.device)
        params, saved, bufs = self._snapshot()
        had_state = len(self.optimizer.state) > 0
        state_saved = None
        if had_state:
            state_saved = [{k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                           for st in self.optimizer.state.values()]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.optimizer.zero_grad(set_to_none=True)
                self._step(static_j)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._restore(params, saved, bufs)
        self.status.zero_()                       # warm-up ran on dummy indices
        if had_state:
            with torch.no_grad():
                for st, old in zip(self.optimizer.state.values(), state_saved):
                    for k, v in old.items():
                        if torch.is_tensor(v):
                            st[k].copy_(v)
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with graph_capture.capture(graph):
            static_loss = self._step(static_j)
        # the capture itself does not execute; state is untouched
        self._keepalive.extend(self.mmb_ops.cached_inv_norms())   # buffers whose addresses the graph holds
        return graph, static_j, static_loss

    def __call__(self, j):
        """Run one step on the row indices j (int64 CUDA tensor); returns the (device) loss."""
        prev = self.mmb_ops.set_status_sink(self.status)
        try:
            n = int(j.shape[0])
            if n not in self.graphs:
                self.graphs[n] = self._capture(n)
            graph, static_j, static_loss = self.graphs[n]
            static_j.copy_(j)
            graph.replay()
            return static_loss
        finally:
            self.mmb_ops.set_status_sink(prev)

    def _capture_epoch(self, sizes):
        """Capture a WHOLE epoch -- one step per batch, the batches being static slices of one index buffer
        -- as a single graph: one replay (and one small index upload) per epoch instead of one per step."""
        t0 = time.perf_counter()
        try:
            return self._capture_epoch_impl(sizes)
        finally:
            CAPTURE_SECONDS[0] += time.perf_counter() - t0

    def _capture_epoch_impl(self, sizes):
        n_total = int(sum(sizes))
        static_flat = torch.zeros(n_total, dtype=torch.int64, device=self.device)
        params, saved, bufs = self._snapshot()
        had_state = len(self.optimizer.state) > 0
        state_saved = None
        if had_state:
            state_saved = [{k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                           for st in self.optimizer.state.values()]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for n in sorted(set(sizes)) * 2:          # warm every batch size (allocator, lazy inits)
                self.optimizer.zero_grad(set_to_none=True)
                self._step(static_flat[:n], self.fork_epoch)
        torch.cuda.current_stream(self.device).wait_stream(side)
        self._restore(params, saved, bufs)
        self.status.zero_()
        if had_state:
            with torch.no_grad():
                for st, old in zip(self.optimizer.state.values(), state_saved):
                    for k, v in old.items():
                        if torch.is_tensor(v):
                            st[k].copy_(v)
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with graph_capture.capture(graph):
            total = torch.zeros((), device=self.device)
            off = 0
            for n in sizes:
                self.optimizer.zero_grad(set_to_none=True)
                total = total + self._step(static_flat[off:off + n], self.fork_epoch)
                off += n
        self._keepalive.extend(self.mmb_ops.cached_inv_norms())
        return graph, static_flat, total

    def run_epoch(self, flat, sizes):
        """All steps of one epoch: ``flat`` = the epoch's row indices in visiting order (int64 CUDA tensor),
        ``sizes`` = the batch sizes.  Returns the (device) sum of the per-step losses, as the reference's
        ``epoch_loss`` accumulates them (simplesif.py:139)."""
        prev = self.mmb_ops.set_status_sink(self.status)
        try:
            key = tuple(int(n) for n in sizes)
            if key not in self.epoch_graphs:
                self.epoch_graphs[key] = self._capture_epoch(key)
            graph, static_flat, total = self.epoch_graphs[key]
            static_flat.copy_(flat)
            graph.replay()
            return total
        finally:
            self.mmb_ops.set_status_sink(prev)

    def check(self):
        from losses import check_status_sink
        check_status_sink(self.status, list(self.gen_model.embed2out.keys()))


class _EpochLosses(list):
    """Per-epoch loss values that may still live on the device: the graph-captured loops append the epoch's
    device scalar (a clone: replays overwrite the graph's static output) and nothing waits for it; ``floats()``
    reads everything back with ONE synchronisation.  The reference reads the value every step
    (simplesif.py:139 ``float(avg_log_prob)``), which stalls the stream each time."""

    def floats(self):
        if not self:
            return []
        if any(torch.is_tensor(x) for x in self):
            vals = torch.stack([x if torch.is_tensor(x) else torch.tensor(float(x), device=self._dev()) for x in self])
            return [float(v) for v in vals.cpu()]
        return [float(x) for x in self]

    def _dev(self):
        for x in self:
            if torch.is_tensor(x):
                return x.device
        return 'cpu'

    def last(self):
        return float(self[-1])


def _epoch_index_batches(dataloader, device):
    """The index batches the DataLoader would produce this epoch, drawing from the same RNG in
    the same order (num_workers = 0: one base-seed draw when the iterator is built, then the
    sampler's own draws), without collating the data sample by sample."""
    torch.empty((), dtype=torch.int64).random_(generator=dataloader.generator)
    batches = list(dataloader.batch_sampler)
    if not batches:
        return
    # one host-to-device copy per epoch, not one per batch
    flat = torch.tensor([i for b in batches for i in b], dtype=torch.int64).to(device, non_blocking=True)
    off = 0
    for b in batches:
        yield flat[off:off + len(b)]
        off += len(b)


def _epoch_indices(dataloader, device):
    """``_epoch_index_batches`` as one flat index tensor + the batch sizes (same draws from the generator)."""
    torch.empty((), dtype=torch.int64).random_(generator=dataloader.generator)
    batches = list(dataloader.batch_sampler)
    flat = torch.tensor([i for b in batches for i in b], dtype=torch.int64).to(device, non_blocking=True)
    return flat, [len(b) for b in batches]


def _graph_epochs(args):
    """Whole-epoch graphs unless ``args['cuda_graph'] == 'step'`` (one graph per step)."""
    return str(args.get('cuda_graph', os.environ.get('MMB_CUDA_GRAPH', '0'))) != 'step'


def _use_cuda_graph(args, gen_model, device):
    flag = args.get('cuda_graph', os.environ.get('MMB_CUDA_GRAPH', '0'))
    if str(flag) in ('0', 'False', 'false', ''):
        return False
    # BatchNorm1d (training mode: batch statistics, running buffers and the step counter updated on the
    # device) is captured like everything else; GraphedStep restores the buffers after its warm-up steps
    return torch.device(device).type == 'cuda'


def optimize_latents(args, train: bool, gen_model, embed_arr, dataloader, n_epochs, lr, word_prob_fn,
                     device, validation_data=None, verbose=True):
    """reference simplesif.py:49-162.

    Minimises ``mean_b(-log p(text, audio, visual | latent_b))`` over the latent matrix (and
    the generator heads when ``train`` and not ``args['freeze_weights']``) with SGD/Adam.
    Returns ``(embeddings (N, d) float32 on device, (losses, all_valid_losses))``.

    ``args['cuda_graph']`` (or MMB_CUDA_GRAPH=1) replays each step as one captured CUDA graph
    (``GraphedStep``); the default is the eager loop, launch for launch what the reference does.
    """
    graphed = _use_cuda_graph(args, gen_model, device)
    train_heads = bool(train and not args['freeze_weights'])
    cache = args.get('_step_cache') if graphed else None
    cache_key = ('latents', train_heads, id(gen_model), id(dataloader.dataset), args['optimizer'], float(lr),
                 str(args.get('cuda_graph')), 'word_loss_weight' in args)
    if cache is not None and cache_key in cache:
        # a captured step for exactly this structure (same generator object, dataset, optimizer, step size) exists:
        # re-use its graphs on fresh latents (sweep.py; the nested validation passes of the e2e loop hit this too)
        stepper, optimizer, embeddings = cache[cache_key]
        stepper.rebind(args, embed_arr)
    else:
        embeddings = torch.tensor(np.array(embed_arr, copy=True), device=device, dtype=torch.float32)
        embeddings.requires_grad = True
        grad_params = [embeddings]
        if train_heads:
            grad_params.extend(gen_model.parameters())
        if graphed and args['optimizer'] == 'adam':
            optimizer = optim.Adam(grad_params, lr=lr, capturable=True)
        else:
            optimizer = _make_optimizer(args, grad_params, lr)
        stepper = GraphedStep(args, gen_model, embeddings, dataloader.dataset, optimizer, word_prob_fn,
                              device) if graphed else None
        if cache is not None:
            cache[cache_key] = (stepper, optimizer, embeddings)
    moments = None if graphed else _dataset_moments(args, dataloader.dataset)

    valid_niter = 10
    start_time = time.time()
    losses = _EpochLosses()
    all_valid_losses = []
    for i in range(n_epochs):
        epoch_loss = torch.zeros((), device=device)
        iters = 0
        if graphed and _graph_epochs(args):
            flat, sizes = _epoch_indices(dataloader, device)
            iters = len(sizes)
            if iters:
                epoch_loss = stepper.run_epoch(flat, sizes).clone()
            if i % valid_niter == 0 or i == n_epochs - 1:
                stepper.check()      # one status read per 10 epochs (and at the end), not one per epoch
        elif graphed:
            for j in _epoch_index_batches(dataloader, device):
                iters += 1
                epoch_loss += stepper(j)
            if i % valid_niter == 0 or i == n_epochs - 1:
                stepper.check()
        else:
            for x in dataloader:
                j, batch_data, batch_masks = _batch_dicts(args, _with_moments(args, x, moments),
                                                          getattr(dataloader.dataset, 'table', None))
                iters += 1
                optimizer.zero_grad()
                out = gen_model(embeddings[j])
                if verbose and i % valid_niter == 0 and iters == 1:
                    _warn_tiny_sigma(out)
                log_prob = -get_log_prob_matrix(args, embeddings[j], out, batch_data, batch_masks, word_prob_fn,
                                                device=device, verbose=False)
                avg_log_prob = log_prob.mean()
                avg_log_prob.backward()
                optimizer.step()
                epoch_loss += avg_log_prob.detach()
        losses.append(epoch_loss if graphed else float(epoch_loss))
        if i % valid_niter == 0:
            if verbose:
                print("epoch {}: {} ({}s)".format(i, losses.last() / max(iters, 1), time.time() - start_time))
            if validation_data is not None and i % (valid_niter * 8) == 0:
                valid_embedding, valid_dataloader = validation_data
                _, valid_losses = optimize_latents(args, False, gen_model, valid_embedding, valid_dataloader,
                                                   n_epochs, lr, word_prob_fn, device, verbose=False)
                print("Validation loss:", valid_losses[0][-1])
                all_valid_losses.append(valid_losses[0][-1])

    if validation_data is not None:  # final validation
        valid_embedding, valid_dataloader = validation_data
        _, valid_losses = optimize_latents(args, False, gen_model, valid_embedding, valid_dataloader,
                                           n_epochs, lr, word_prob_fn, device, verbose=False)
        print("(Final) Validation loss:", valid_losses[0][-1])
        all_valid_losses.append(valid_losses[0][-1])

    if cache is not None:
        return embeddings.detach().clone(), (losses.floats(), all_valid_losses)    # the static tensor serves the next call
    embeddings.requires_grad = False
    return embeddings, (losses.floats(), all_valid_losses)


def make_word_log_prob_fn(args, weights, word_embeddings, a=1e-3):
    """The ``get_word_log_prob2`` closure of reference simplesif.py:505-537."""
    if args['word_sim_metric'] == 'angular':
        word_log_prob_fn = get_word_log_prob_angular2
    elif args['word_sim_metric'] == 'dot_prod':
        word_log_prob_fn = get_word_log_prob_dot_prod
    else:
        raise NotImplementedError

    def get_word_log_prob2(latents, word_weights, sent_embeddings, mask):
        if word_log_prob_fn is not get_word_log_prob_angular2:
            return word_log_prob_fn(latents, word_embeddings, word_weights, sent_embeddings, mask, a)
        # the inf check of reference 529-535: the kernel raises a bit in `status`; read here (one sync,
        # as in the reference) unless a status sink defers it to the epoch's end (graph-captured loop)
        status = mmb_ops.new_status(latents.device)
        word_log_prob = word_log_prob_fn(latents, word_embeddings, word_weights, sent_embeddings, mask, a,
                                         status=status)
        if not mmb_ops.status_deferred() and int(status.item()) & 2:
            print('word inf')
            print(latents.size())
            sys.exit()
        return word_log_prob
    return get_word_log_prob2


def train_end_to_end(args, gen_model, senti_model, train_embedding, dataloader, senti_train_data, senti_mask,
                     word_prob_fn, device, n_epochs=None, verbose=True, validation_data=None):
    """The e2e loop of reference simplesif.py:694-800: latents, generator heads and the
    sentiment regressor are optimised jointly on
    ``likelihood_weight * (-log p) + (1 - likelihood_weight) * L1(sentiment) * senti_mask``.
    ``validation_data = (valid_embedding, valid_dataloader)``: every 80 epochs the validation latents
    are optimised for ``n_epochs`` with the current heads and their last epoch loss recorded
    (reference 796-800) -- besides the number it reports, that pass draws from torch's global
    generator (one DataLoader base seed per epoch), so it is part of reproducing the reference's
    batch order.  Returns ``(train_embed, (train_losses, all_valid_losses))``."""
    graphed = _use_cuda_graph(args, gen_model, device)
    cache = args.get('_step_cache') if graphed else None
    cache_key = ('e2e', id(gen_model), id(senti_model), id(senti_train_data), id(senti_mask), id(dataloader.dataset),
                 args['optimizer'], float(args['lr']), str(args.get('cuda_graph')), 'word_loss_weight' in args)
    loss_function = nn.L1Loss(reduction='none')
    like_w = args['likelihood_weight']
    holder = {}

    def mixed_loss(j, e, log_prob):
        """reference simplesif.py:776-786: L1 sentiment term, masked, mixed with the likelihood."""
        _, s_data = senti_train_data[j]
        senti_loss = loss_function(senti_model(e), s_data)
        if senti_loss.dim() > 1:
            senti_loss = senti_loss.mean(-1)
        senti_loss = senti_loss * senti_mask[j].reshape(senti_loss.shape)
        st = holder.get('stepper')
        if st is not None and st.cache_mode:                      # weights from device memory (GraphedStep.rebind)
            return st.loss_w[2] * log_prob + st.loss_w[3] * senti_loss
        return like_w * log_prob + (1. - like_w) * senti_loss

    if cache is not None and cache_key in cache:
        stepper, optimizer, train_embed = cache[cache_key]
        stepper.rebind(args, train_embedding)
    else:
        train_embed = torch.tensor(np.array(train_embedding, copy=True), device=device, dtype=torch.float32)
        train_embed.requires_grad = True
        grad_params = [train_embed] + list(gen_model.parameters()) + list(senti_model.parameters())
        if graphed and args['optimizer'] == 'adam':
            optimizer = optim.Adam(grad_params, lr=args['lr'], capturable=True)
        else:
            optimizer = _make_optimizer(args, grad_params, args['lr'])
        stepper = GraphedStep(args, gen_model, train_embed, dataloader.dataset, optimizer, word_prob_fn, device,
                              extra_loss=mixed_loss, extra_modules=[senti_model]) if graphed else None
        holder['stepper'] = stepper
        if cache is not None:
            cache[cache_key] = (stepper, optimizer, train_embed)
    moments = None if graphed else _dataset_moments(args, dataloader.dataset)
    train_losses, all_valid_losses = _EpochLosses(), []
    start_time = time.time()
    n_epochs = n_epochs if n_epochs is not None else args['n_epochs']
    for i in range(n_epochs):
        epoch_loss = torch.zeros((), device=device)
        iters = 0
        if graphed and _graph_epochs(args):
            flat, sizes = _epoch_indices(dataloader, device)
            iters = len(sizes)
            if iters:
                epoch_loss = stepper.run_epoch(flat, sizes).clone()
            if i % 10 == 0 or i == n_epochs - 1:
                stepper.check()
        elif graphed:
            for j in _epoch_index_batches(dataloader, device):
                iters += 1
                epoch_loss += stepper(j)
            if i % 10 == 0 or i == n_epochs - 1:
                stepper.check()
        else:
            for x in dataloader:
                j, batch_data, batch_masks = _batch_dicts(args, _with_moments(args, x, moments),
                                                          getattr(dataloader.dataset, 'table', None))
                iters += 1
                optimizer.zero_grad()
                out = gen_model(train_embed[j])
                log_prob = -get_log_prob_matrix(args, train_embed[j], out, batch_data, batch_masks, word_prob_fn,
                                                device=device, verbose=False)
                loss = mixed_loss(j, train_embed[j], log_prob).mean()
                loss.backward()
                optimizer.step()
                epoch_loss += loss.detach()
        train_losses.append(epoch_loss if graphed else float(epoch_loss))
        if i % 10 == 0:
            if verbose:
                print("epoch {}: {} ({}s)".format(i, train_losses.last() / max(iters, 1), time.time() - start_time))
            if validation_data is not None and i % 80 == 0:
                valid_embedding, valid_dataloader = validation_data
                _, (valid_losses, _) = optimize_latents(args, False, gen_model, valid_embedding, valid_dataloader,
                                                        n_epochs, args['lr'], word_prob_fn, device, verbose=False)
                if verbose:
                    print("Validation loss:", valid_losses[-1])
                all_valid_losses.append(valid_losses[-1])
    if cache is not None:
        return train_embed.detach().clone(), (train_losses.floats(), all_valid_losses)
    train_embed.requires_grad = False
    return train_embed, (train_losses.floats(), all_valid_losses)


def read_config(config_file):
    """reference simplesif.py:177-184."""
    config = json.load(open(config_file, 'r'))
    pprint.PrettyPrinter(indent=2).pprint(config)
    return config


def parse_arguments(argv=None):
    """reference simplesif.py:186-238 -- same flags; the JSON config is merged over them and
    ``--pos_embed_dim`` / ``--e2e`` / ``--sentiment_epochs`` override it."""
    parser = argparse.ArgumentParser()
    parser.add_argument('config_file', help='JSON file containing hyperparameters for model')
    parser.add_argument('dataset', choices=['mosi', 'pom', 'iemocap'])
    parser.add_argument('--unimodal', action='store_true', help='run mmb1 (unimodal factorization)')
    parser.add_argument('--pos_embed_dim', type=int)
    parser.add_argument('--batch_size', type=int, default=64)
    parser.add_argument('--n_runs', type=int, default=1)
    parser.add_argument('--semi_sup_idxes', choices=['{:.1f}'.format(x) for x in np.arange(0.1, 1, 0.1)])
    parser.add_argument('--config_name', help='override config name in config file')
    parser.add_argument('--lr_decay', type=float, default=0.5)
    parser.add_argument('--early_stopping', action='store_true',
                        help='early stopping when training sentiment model')
    parser.add_argument('--sentiment_epochs', type=int)
    parser.add_argument('--emotion', choices=['happy', 'angry', 'neutral', 'sad'], help='iemocap emotion')
    parser.add_argument('--optimizer', choices=['sgd', 'adam'], default='sgd')
    parser.add_argument('--norm', choices=['layer_norm', 'batch_norm'])
    parser.add_argument('--likelihood_weight', type=float)
    parser.add_argument('--e2e', choices=['y', 'n'], help='end-to-end training of latent variables')
    parser.add_argument('--time_test', action='store_true', help='Run inference timing')
    parser.add_argument('--cuda_device', type=int, choices=list(range(8)), help='set CUDA device number')
    parser.add_argument('--cuda', action='store_true')
    args = vars(parser.parse_args(argv))

    override_dict = {}
    if args['pos_embed_dim'] is not None:
        override_dict['pos_embed_dim'] = args['pos_embed_dim']
    if args['e2e'] is not None:
        override_dict['e2e'] = args['e2e']
    config = read_config(args['config_file'])
    print('######################################')
    print("Config: {}".format(config['config_num']))
    args.update(config)
    args.update(override_dict)
    if args['e2e'] == 'y':
        args['e2e'] = True
    elif args['e2e'] == 'n':
        args['e2e'] = False
    if args['sentiment_epochs']:
        args['n_sentiment_epochs'] = args['sentiment_epochs']
    return args


def prepare_splits(args, word_embeddings, weights, splits, masks, device):
    """reference simplesif.py:296-399: SIF initialisation per split (PC removed per split),
    id -> vector expansion, positional features on audio / visual."""
    id_key = 'text' if args['dataset'] == 'mosi' else 'text_id'
    embeddings = [get_sentence_embeddings(word_embeddings, weights, s[id_key]) for s in splits]
    w_t = torch.tensor(weights, device=device, dtype=torch.float32)
    we_t = torch.tensor(word_embeddings, device=device, dtype=torch.float32)
    for s, m in zip(splits, masks):
        ids = torch.as_tensor(s[id_key], device=device)
        if args['dataset'] == 'mosi':
            s['text_id'] = s['text']
        else:
            s['text_align'] = s['text']
            update_masks_vect(m, s['text_align'], 'text_align')
        if args.get('text_ids'):
            # keep the transcript as ids; run_experiment builds MMDataIds from them
            s['text'] = TokenIds(ids, we_t)
            s['text_weights'] = w_t
        else:
            s['text'] = we_t[ids]
            s['text_weights'] = w_t[ids]
        if args.get('pos_embed_dim', 0) and args['pos_embed_dim'] > 0:
            n_points, seq_len = m['covarep'].shape[:2]
            ext = np.ones((n_points, seq_len, args['pos_embed_dim']), dtype=np.int64)
            for k in ('covarep', 'facet'):
                s[k] = add_positional_embeddings(args, s[k])
                m[k] = np.concatenate([m[k], ext], axis=-1)
    return embeddings, w_t, we_t


def run_experiment(args, word_embeddings, weights, splits, masks, device, folder=None):
    """Everything reference simplesif.py:294-914 does once the splits are loaded and normalised
    (``masks`` already hold the text mask of ``update_masks``): SIF initialisation per split, id
    expansion and positional columns, datasets / loaders, then per run the e2e branch (625-806) or
    the two-stage branch (541-624), latent inference for valid / test and the downstream regressor.
    ``folder`` (a format string taking the run number, or None) is where the reference's artefacts
    go.  Returns the list of per-run ``(results, train_losses, (train, valid, test) latents)``."""
    embeddings, w_t, we_t = prepare_splits(args, word_embeddings, weights, splits, masks, device)
    train, valid, test = splits

    def dataset(s, m):
        if isinstance(s['text'], TokenIds):
            t = s['text']
            if args['dataset'] == 'mosi':
                return MMDataIds(t.ids, s['covarep'], s['facet'], m, s['text_weights'], t.table, device)
            return MMDataExtraIds(t.ids, s['covarep'], s['facet'], m, s['text_weights'], t.table, s['text_align'],
                                  device)
        if args['dataset'] == 'mosi':
            return MMData(s['text'], s['covarep'], s['facet'], m, s['text_weights'], device)
        return MMDataExtra(s['text'], s['covarep'], s['facet'], m, s['text_weights'], s['text_align'], device)
    bs = args['batch_size']
    loaders = [DataLoader(dataset(train, masks[0]), batch_size=bs, shuffle=True),
               DataLoader(dataset(valid, masks[1]), batch_size=bs * 8),
               DataLoader(dataset(test, masks[2]), batch_size=bs * 8)]
    word_fn = make_word_log_prob_fn(args, w_t, we_t)
    d, A, Vd = train['text'].shape[-1], train['covarep'].shape[-1], train['facet'].shape[-1]
    sentiment_data = (train['label'], valid['label'], test['label'])

    runs = []
    for r in range(args['n_runs']):
        run_dir = folder.format(r) if folder is not None else None
        if run_dir is not None:
            for sub in ('pre', 'post'):
                os.makedirs(os.path.join(run_dir, sub), exist_ok=True)
            json.dump(args, open(os.path.join(run_dir, 'config.json'), 'w'), indent=2)
            torch.save(torch.tensor(np.concatenate(embeddings, axis=0), device=device, dtype=torch.float32),
                       os.path.join(run_dir, 'pre', 'embed.bin'))
        gen_model = AudioVisualGeneratorMultimodal(d, A, Vd, norm=args['norm'], frozen_weights=args['freeze_weights'],
                                                   unimodal=args['unimodal']).to(device)
        n_epochs, lr = args['n_epochs'], args['lr']
        if args['e2e']:
            n_out = 1 if train['label'].ndim == 1 else train['label'].shape[-1]
            senti_model = SentimentModel(d, args['sentiment_hidden_size'], n_out).to(device)
            senti_mask = torch.ones(len(train['label']), device=device)
            train_embed, (train_losses, valid_losses) = train_end_to_end(
                args, gen_model, senti_model, embeddings[0], loaders[0], SentimentData(train['label'], device),
                senti_mask, word_fn, device, validation_data=(embeddings[1], loaders[1]))
        else:
            train_embed, (train_losses, valid_losses) = optimize_latents(
                args, True, gen_model, embeddings[0], loaders[0], n_epochs, lr, word_fn, device,
                validation_data=(embeddings[1], loaders[1]))
        valid_embed, _ = optimize_latents(args, False, gen_model, embeddings[1], loaders[1], n_epochs, lr,
                                          word_fn, device)
        test_embed, (test_losses, _) = optimize_latents(args, False, gen_model, embeddings[2], loaders[2], n_epochs,
                                                        lr, word_fn, device)
        post = None
        if run_dir is not None:
            for name, vals in (('embed_loss.txt', train_losses), ('embed_valid_loss.txt', valid_losses),
                               ('embed_test_loss.txt', test_losses)):
                with open(os.path.join(run_dir, name), 'w') as f:
                    f.writelines('{}\n'.format(v) for v in vals)
            post = os.path.join(run_dir, 'post')
            torch.save(torch.cat([train_embed, valid_embed, test_embed], dim=0), os.path.join(post, 'embed.bin'))
        results, _ = train_sentiment_for_latents(args, (train_embed, valid_embed, test_embed), sentiment_data, device,
                                                 model_save_path=post)
        runs.append((results, train_losses, (train_embed, valid_embed, test_embed)))
    return runs


def main(argv=None):
    """reference simplesif.py:240-916, for the datasets the reference's loaders can open."""
    args = parse_arguments(argv)
    if args['cuda_device'] is not None:
        os.environ['CUDA_VISIBLE_DEVICES'] = str(args['cuda_device'])
    if not torch.cuda.is_available():
        raise RuntimeError('this implementation runs on a CUDA device (sm_100a) only')
    device = torch.device('cuda')

    word2ix, word_embeddings, data = load_data(args)
    splits = list(data)
    masks = []
    for k in range(3):
        splits[k], m = normalize_data(splits[k])
        update_masks(m, splits[k]['text' if args['dataset'] == 'mosi' else 'text_id'], word_embeddings.shape[-1])
        masks.append(m)
    weights = load_weights(args)
    if args['word_sim_metric'] == 'dot_prod':
        word_embeddings = word_embeddings / np.linalg.norm(word_embeddings, axis=-1, keepdims=True)
    config_name = args['config_name'] or os.path.split(os.path.split(args['config_file'])[0])[1]
    folder = 'model_saves/{}/config_{}_run_{{}}'.format(config_name, args['config_num'])
    run_experiment(args, word_embeddings, weights, splits, masks, device, folder=folder)


if __name__ == '__main__':
    main()
