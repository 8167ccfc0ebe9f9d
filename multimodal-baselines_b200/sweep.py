"""The reference's hyper-parameter grid evaluated in one pass (BASELINE.json configs[4]; SURVEY.md
§8d "Config 5", §8e "Sweep": replicas only -- the 512 configs are independent, so they shard over
the GPUs of the box with no data-path collective).

``make_grid()`` is the grid of reference configs/make_configs.py:16-32 (2^9 = 512 combinations
of sentiment_hidden_size, lr, sentiment_lr, n_epochs, word_loss_weight, likelihood_weight,
pos_embed_dim, norm, optimizer) in ``itertools.product`` order -- the reference shuffles the
list with an unseeded ``random.shuffle`` (line 53), so its ``config_<i>.json`` numbering is
different on every generation; here ``config_num`` is the product index.

``run_config`` follows the e2e branch of reference simplesif.py:625-914 for one config on
already-prepared splits: end-to-end training of latents + generator heads + sentiment regressor,
latent inference for valid / test, then ``train_sentiment_for_latents``.  The SIF initialisation
(reference 296-311) does not depend on the config and is computed once per process.  Every step
is a CUDA-graph replay where the norm allows it (LayerNorm; BatchNorm1d configs run eagerly).

    python sweep.py [--limit K] [--epochs-scale S] [--out results.jsonl]
    torchrun --nproc-per-node 8 sweep.py ...        # configs rank, rank + world, ...
"""
import argparse
import itertools
import warnings
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

GRID = {   # reference configs/make_configs.py:16-32
    'sentiment_hidden_size': [100, 150],
    'lr': [1e-3, 1e-4],
    'sentiment_lr': [1e-1, 1e-2],
    'seq_len': [20],
    'word_sim_metric': ['angular'],
    'n_epochs': [100, 200],
    'freeze_weights': [False],
    'n_sentiment_epochs': [400],
    'word_loss_weight': [0.001, 0.002],
    'likelihood_weight': [0.0001, 0.001],
    'pos_embed_dim': [2, 4],
    'e2e': [True],
    'norm': ['layer_norm', 'batch_norm'],
    'optimizer': ['sgd', 'adam'],
}


def make_grid():
    keys = list(GRID.keys())
    configs = []
    for i, values in enumerate(itertools.product(*[GRID[k] for k in keys])):
        cfg = dict(zip(keys, values))
        cfg['config_num'] = i
        configs.append(cfg)
    return configs


def synthetic_mosi(seed=0, sizes=(1284, 229, 686), L=20, V=3016, d=300, A=74, Vd=47):
    """MOSI-shaped splits (SURVEY.md §8d Config 2): Zipf ids right-padded with 0, raw audio / visual
    features with exact zeros on the padded steps (what normalize_data keys on), and a sentiment
    label that is a noisy linear read-out of the utterance's mean word vector and mean audio."""
    rng = np.random.default_rng(seed)
    We = (0.4 * rng.standard_normal((V, d)) + 0.3 * rng.standard_normal((1, d))).astype(np.float32)
    We[0] = 0
    p = 1.0 / np.arange(1, V) ** 1.1
    p /= p.sum()
    weights = np.concatenate([[1.0], 1e-3 / (1e-3 + p)])
    w_txt, w_aud = rng.standard_normal(d) / np.sqrt(d), rng.standard_normal(A) / np.sqrt(A)
    splits = []
    for n in sizes:
        ids = rng.choice(np.arange(1, V), size=(n, L), p=p).astype(np.int64)
        lens = rng.integers(1, L + 1, size=n)
        pad = np.arange(L)[None, :] >= lens[:, None]
        ids[pad] = 0
        cov = (rng.standard_normal((n, L, A)) * 2 + 1).astype(np.float32)
        fac = (rng.standard_normal((n, L, Vd)) * 1.5 - 0.5).astype(np.float32)
        cov[pad] = 0
        fac[pad] = 0
        mean_txt = We[ids].sum(1) / lens[:, None]
        mean_aud = cov.sum(1) / lens[:, None]
        label = 3 * np.tanh(mean_txt @ w_txt * 4 + mean_aud @ w_aud) + 0.3 * rng.standard_normal(n)
        splits.append({'text': ids, 'covarep': cov, 'facet': fac, 'label': label.astype(np.float32)})
    return We, weights, splits


class Prepared(object):
    """Per-process state shared by all configs: SIF embeddings per split and, per
    ``pos_embed_dim``, the device datasets."""

    def __init__(self, We, weights, splits, device, batch_size=64):
        import copy
        import simplesif
        import utils
        self.device, self.batch_size = device, batch_size
        self.labels = [s['label'] for s in splits]
        raw = [copy.deepcopy(s) for s in splits]
        masks = []
        for k in range(3):
            raw[k], m = utils.normalize_data(raw[k])
            simplesif.update_masks(m, raw[k]['text'], We.shape[-1])
            masks.append(m)
        self.raw, self.masks, self.We, self.weights = raw, masks, We, weights
        self.by_pos = {}
        self.embeddings = None
        # Captured steps and the modules they were captured on, kept across grid points (``graph_cache``): grid
        # points of the same structure (pos_embed_dim, norm, optimizer, step size, regressor width) re-use the graphs;
        # their parameters are re-initialised IN PLACE from a freshly constructed module (same draws from the
        # generator, same values), their likelihood weights are device scalars (simplesif.GraphedStep.rebind).
        self.step_cache = {}
        self.modules = {}
        self.senti_train_data = None
        self.senti_mask = None

    def persistent_module(self, key, fresh):
        """The module kept for ``key``, holding ``fresh``'s parameters and buffers (copied in place)."""
        if key not in self.modules:
            self.modules[key] = fresh.to(self.device)
        else:
            self.modules[key].load_state_dict(fresh.state_dict())
        return self.modules[key]

    def for_pos(self, pos_embed_dim):
        import copy
        import simplesif
        import utils
        from torch.utils.data import DataLoader
        if pos_embed_dim not in self.by_pos:
            splits, masks = copy.deepcopy(self.raw), copy.deepcopy(self.masks)
            args = {'dataset': 'mosi', 'pos_embed_dim': pos_embed_dim}
            emb, w_t, we_t = simplesif.prepare_splits(args, self.We, self.weights, splits, masks, self.device)
            if self.embeddings is None:
                self.embeddings = emb          # SIF embeddings do not depend on the config
            ds = [utils.MMData(s['text'], s['covarep'], s['facet'], m, s['text_weights'], self.device)
                  for s, m in zip(splits, masks)]
            bs = self.batch_size
            loaders = [DataLoader(ds[0], batch_size=bs, shuffle=True), DataLoader(ds[1], batch_size=bs * 8),
                       DataLoader(ds[2], batch_size=bs * 8)]
            dims = (splits[0]['text'].shape[-1], splits[0]['covarep'].shape[-1], splits[0]['facet'].shape[-1])
            self.by_pos[pos_embed_dim] = (loaders, w_t, we_t, dims)
        return self.by_pos[pos_embed_dim]


def run_config(cfg, prep, epochs_scale=1.0, cuda_graph=True, verbose=False, defer_regressor=False, graph_cache=False):
    """One grid point, reference simplesif.py:625-914 (e2e branch).  Returns the test metrics of
    the downstream regressor and the final losses.  ``defer_regressor``: stop after the latent phases and
    return the regressor problem as a ``sentiment_batched.RegressorJob`` under ``'job'`` (with the random
    stream captured where the regressor would start), to be trained together with other grid points'."""
    import simplesif
    from models import AudioVisualGeneratorMultimodal
    from sentiment_model import SentimentData, SentimentModel, train_sentiment_for_latents
    device = prep.device
    args = {'dataset': 'mosi', 'unimodal': False, 'early_stopping': False, 'lr_decay': 0.5,
            'cuda_graph': 1 if cuda_graph else 0}
    args.update(cfg)
    args['n_epochs'] = max(1, int(round(cfg['n_epochs'] * epochs_scale)))
    args['n_sentiment_epochs'] = max(1, int(round(cfg['n_sentiment_epochs'] * epochs_scale)))
    loaders, w_t, we_t, (d, A, Vd) = prep.for_pos(cfg['pos_embed_dim'])
    torch.manual_seed(1000 + cfg['config_num'])
    word_fn = simplesif.make_word_log_prob_fn(args, w_t, we_t)
    gen_model = AudioVisualGeneratorMultimodal(d, A, Vd, norm=args['norm'], frozen_weights=args['freeze_weights'],
                                               unimodal=False)
    senti_model = SentimentModel(d, args['sentiment_hidden_size'], 1)
    if graph_cache and cuda_graph:
        # ``graph_cache``: same values as a fresh module, but in the persistent module the cached graphs read
        args['_step_cache'] = prep.step_cache
        gen_model = prep.persistent_module(('gen', cfg['pos_embed_dim'], args['norm'], args['freeze_weights']), gen_model)
        senti_model = prep.persistent_module(('senti', args['sentiment_hidden_size']), senti_model)
        if prep.senti_train_data is None:
            prep.senti_train_data = SentimentData(prep.labels[0], device)
            prep.senti_mask = torch.ones(len(prep.labels[0]), device=device)
        senti_train_data, senti_mask = prep.senti_train_data, prep.senti_mask
    else:
        gen_model, senti_model = gen_model.to(device), senti_model.to(device)
        senti_train_data = SentimentData(prep.labels[0], device)
        senti_mask = torch.ones(len(prep.labels[0]), device=device)
    # The reference's helpers print progress and, on a non-finite log-probability, print the
    # modality names and sys.exit() (losses.py:258-264) -- in the reference one config is one
    # process; here that exit marks the config as diverged and the sweep goes on.
    import io
    buf = io.StringIO()
    old_stdout = sys.stdout
    if not verbose:
        sys.stdout = buf
    phase, t_last = {}, [time.perf_counter()]
    cap0 = simplesif.CAPTURE_SECONDS[0]

    def lap(name):
        torch.cuda.synchronize(device)
        now = time.perf_counter()
        phase[name] = now - t_last[0]
        t_last[0] = now
    try:
        train_embed, (train_losses, _) = simplesif.train_end_to_end(
            args, gen_model, senti_model, prep.embeddings[0], loaders[0], senti_train_data,
            senti_mask, word_fn, device, verbose=False, validation_data=(prep.embeddings[1], loaders[1]))
        lap('train_e2e_incl_nested_validation')
        valid_embed, _ = simplesif.optimize_latents(args, False, gen_model, prep.embeddings[1], loaders[1],
                                                    args['n_epochs'], args['lr'], word_fn, device, verbose=False)
        test_embed, (test_losses, _) = simplesif.optimize_latents(args, False, gen_model, prep.embeddings[2],
                                                                  loaders[2], args['n_epochs'], args['lr'], word_fn,
                                                                  device, verbose=False)
        lap('valid_test_latents')
        if defer_regressor:
            import sentiment_batched
            if sentiment_batched.can_batch(args, prep.labels):
                job = sentiment_batched.RegressorJob(args, (train_embed, valid_embed, test_embed), tuple(prep.labels),
                                                     tag=cfg['config_num'])
                phase['graph_capture_latent_loops'] = simplesif.CAPTURE_SECONDS[0] - cap0
                sys.stdout = old_stdout
                return {'config_num': cfg['config_num'], 'job': job, 'train_loss': train_losses[-1],
                        'test_loss': test_losses[-1], 'phase_s': {k: round(v, 3) for k, v in phase.items()}}
        results, _ = train_sentiment_for_latents(args, (train_embed, valid_embed, test_embed), tuple(prep.labels),
                                                 device)
        lap('sentiment_regressor')
        phase['graph_capture_latent_loops'] = simplesif.CAPTURE_SECONDS[0] - cap0
    except SystemExit:
        sys.stdout = old_stdout
        tail = ' | '.join(buf.getvalue().strip().splitlines()[-7:])
        return {'config_num': cfg['config_num'], 'diverged': True, 'message': tail}
    finally:
        sys.stdout = old_stdout
    results = {k: v for k, v in results.items() if k in ('mae', 'accuracy', 'corr', 'mult_acc', 'f_score')}
    return {'config_num': cfg['config_num'], 'results': results, 'train_loss': train_losses[-1],
            'test_loss': test_losses[-1], 'phase_s': {k: round(v, 3) for k, v in phase.items()}}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--limit', type=int, default=0, help='only the first K configs of the grid (0 = all 512)')
    ap.add_argument('--epochs-scale', type=float, default=1.0, help='scale n_epochs / n_sentiment_epochs (smoke runs)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--only', type=int, nargs='*', default=None, help='run just these config numbers')
    ap.add_argument('--out', default='')
    ap.add_argument('--no-graph-cache', action='store_true',
                    help='capture every grid point\'s graphs anew (default: grid points of the same structure '
                         're-use the captured steps; results are bit-identical either way)')
    ap.add_argument('--regressor-batch', type=int, default=8,
                    help='train the downstream regressors of this many grid points as one batched model '
                         '(SURVEY 8f N4; 1 = the sequential module per grid point)')
    a = ap.parse_args(argv)
    import torch.distributed as dist
    world, rank = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise RuntimeError('sweep.py runs on CUDA devices only (no CPU fallback)')
    # More ranks than GPUs is allowed and useful: rank r runs on GPU r % n_gpus, so K = world / n_gpus
    # processes share each GPU.  One config keeps a B200 busy for a fraction of the time (B = 64 kernels
    # between graph captures and host-side bookkeeping); K processes fill each other's gaps.
    n_dev = torch.cuda.device_count()
    local = local % n_dev
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('gloo')          # results are gathered as Python objects; no GPU collective
    warnings.filterwarnings('ignore', message='.*ill-defined.*')   # sklearn's report on empty classes
    grid = make_grid()
    if a.limit:
        grid = grid[:a.limit]
    if a.only:
        grid = [c for c in grid if c['config_num'] in set(a.only)]
    mine = grid[rank::world]
    We, weights, splits = synthetic_mosi()
    devnull = open(os.devnull, 'w')
    old = sys.stdout
    sys.stdout = devnull                        # the reference's helpers print shapes / banners
    try:
        prep = Prepared(We, weights, splits, device)
        prep.for_pos(2)
    finally:
        sys.stdout = old
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = []
    if a.regressor_batch > 1:
        import sentiment_batched
        for o in range(0, len(mine), a.regressor_batch):
            part = [run_config(cfg, prep, a.epochs_scale, not a.no_graph, defer_regressor=True,
                               graph_cache=not a.no_graph_cache) for cfg in mine[o:o + a.regressor_batch]]
            jobs = [r['job'] for r in part if 'job' in r]
            t_reg = time.perf_counter()
            buf, old_stdout = open(os.devnull, 'w'), sys.stdout
            sys.stdout = buf                         # the metric helpers print reports
            try:
                sentiment_batched.run_jobs(jobs, device, max_batch=a.regressor_batch)
            finally:
                sys.stdout = old_stdout
            torch.cuda.synchronize()
            t_reg = (time.perf_counter() - t_reg) / max(len(jobs), 1)
            for r in part:
                job = r.pop('job', None)
                if job is not None:
                    r['results'] = {k: v for k, v in job.results.items()
                                    if k in ('mae', 'accuracy', 'corr', 'mult_acc', 'f_score')}
                    r['phase_s']['sentiment_regressor_batched_share'] = round(t_reg, 3)
            out.extend(part)
    else:
        out = [run_config(cfg, prep, a.epochs_scale, not a.no_graph, graph_cache=not a.no_graph_cache) for cfg in mine]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (out, dt))
        out = [r for part, _ in gathered for r in part]
        dt = max(t for _, t in gathered)
        dist.destroy_process_group()
    if rank == 0:
        out.sort(key=lambda r: r['config_num'])
        if a.out:
            with open(a.out, 'w') as f:
                for r in out:
                    f.write(json.dumps(r) + '\n')
        maes = [r['results']['mae'] for r in out if 'results' in r]
        print(json.dumps({'metric': 'grid configs/s (e2e train + valid/test latents + sentiment regressor)',
                          'configs': len(out), 'n_gpus': min(world, n_dev), 'processes': world,
                          'processes_per_gpu': max(1, world // max(n_dev, 1)), 'seconds': dt, 'value': len(out) / dt,
                          'value_per_gpu': len(out) / dt / max(1, min(world, n_dev)),
                          'epochs_scale': a.epochs_scale, 'cuda_graph': not a.no_graph,
                          'regressor_batch': a.regressor_batch, 'graph_cache': not a.no_graph_cache,
                          'diverged': [r['config_num'] for r in out if r.get('diverged')],
                          'best_test_MAE': (min(maes) if maes else None)}), flush=True)


if __name__ == '__main__':
    main()
