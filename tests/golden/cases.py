"""Seeded input builders shared by ``make_golden.py`` (reference side) and the tests.

Everything is generated from ``np.random.default_rng(seed)`` / ``np.random.RandomState``;
``checksum`` is stored in the fixtures so that a test notices if a NumPy upgrade ever
changes a stream instead of silently comparing against the wrong inputs.
"""
import numpy as np


def checksum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a.ravel() * np.cos(np.arange(a.size))).sum()])


def table(V, d, seed, planted=True):
    """Synthetic GloVe-like table: ``0.4 N(0,1) + mu`` with a fixed common
    direction ``mu = 0.3 N(0,1)`` (real GloVe has one; it gives compute_pc a spectral gap),
    row 0 (pad / OOV row) zero, float32."""
    rng = np.random.default_rng(seed)
    We = 0.4 * rng.standard_normal((V, d))
    if planted:
        We += 0.3 * rng.standard_normal((1, d))
    We[0] = 0.0
    return We.astype(np.float32)


def zipf_ids(rng, n, L, V, lo_len=1, s=1.1):
    """Right-padded (pad id 0) Zipf(s) token ids over 1..V-1, lengths ~ U[lo_len, L]."""
    p = 1.0 / np.arange(1, V, dtype=np.float64) ** s
    p /= p.sum()
    ids = rng.choice(np.arange(1, V), size=(n, L), p=p).astype(np.int64)
    lens = rng.integers(lo_len, L + 1, size=n)
    ids[np.arange(L)[None, :] >= lens[:, None]] = 0
    return ids, p


def sif_weights(p, a=1e-3):
    """sif.py:14-32 form ``a/(a+p(w))``; index 0 (pad/OOV) gets 1.0 like the POM fixture."""
    w = np.empty(p.size + 1, dtype=np.float64)
    w[0] = 1.0
    w[1:] = a / (a + p)
    return w


def sif_mosi_like(n=64, L=20, V=300, d=300, seed=1):
    rng = np.random.default_rng(seed)
    ids, p = zipf_ids(rng, n, L, V)
    return table(V, d, seed + 100), sif_weights(p), ids


def sif_tall(n=330, L=6, V=120, d=300, seed=2):
    """N >= d (no sklearn transpose); a few negative ids (weight 0, sif_functions.py:12)
    and a vocabulary entry whose weight is exactly 0 (does not count in the divisor)."""
    rng = np.random.default_rng(seed)
    ids, p = zipf_ids(rng, n, L, V, lo_len=2)
    w = sif_weights(p)
    w[7] = 0.0
    ids[3, 1] = -1
    ids[5, 0] = -2
    ids[9, 2] = 7
    return table(V, d, seed + 100), w, ids


PC_CASES = {
    # tag: (N, planted-gap?, seed)
    'tall_gap': (350, True, 21),
    'tall_nogap': (350, False, 22),
    'short_gap': (120, True, 23),
    'short_nogap': (40, False, 24),
}


def pc_matrix(n, gap, seed, d=300):
    rng = np.random.default_rng(seed)
    X = 0.1 * rng.standard_normal((n, d))
    if gap:
        X += 0.3 * rng.standard_normal((1, d)) * (1.0 + 0.1 * rng.standard_normal((n, 1)))
    return X


MMB_CASES = {
    'mmb2_small': dict(B=5, T=4, d=24, A=7, Vd=5, V=40, seed=31, norm='layer_norm',
                       unimodal=False, args={'word_loss_weight': 0.3}),
    'mmb1_small': dict(B=4, T=6, d=24, A=6, Vd=9, V=30, seed=32, norm=None,
                       unimodal=True, args={}),
    'mmb2_bn': dict(B=6, T=3, d=16, A=4, Vd=3, V=25, seed=33, norm='batch_norm',
                    unimodal=False, args={'word_loss_weight': 0.5}),
    'mmb2_mosi': dict(B=8, T=20, d=300, A=76, Vd=49, V=200, seed=34, norm='layer_norm',
                      unimodal=False, args={'word_loss_weight': 0.1}),
}


def head_dims(d, A, Vd, unimodal):
    if unimodal:
        return {'audio': A, 'visual': Vd}
    return {'audio': A, 'visual': Vd, 'audiovisual': A + Vd, 'textaudio': d + A,
            'textvisual': d + Vd, 'textaudiovisual': d + A + Vd}


def mmb_inputs(B, T, d, A, Vd, V, seed, norm, unimodal, args):
    """One batch in the layout of ``MMData.__getitem__`` (utils.py:231-233): text (B,T,d)
    word vectors, text mask (B,T,d) broadcast of ``ids != 0`` (simplesif.py:36-40), audio /
    visual in [-1,1] with trailing pad steps = -10 and mask 0 (utils.py:171-189)."""
    rng = np.random.default_rng(seed)
    We = table(V, d, seed + 100)
    ids, p = zipf_ids(rng, B, T, V)
    ww = sif_weights(p).astype(np.float32)
    lens = (ids != 0).sum(1)
    text = We[ids]
    text_m = np.broadcast_to((ids != 0)[:, :, None], (B, T, d)).astype(np.float32).copy()
    step = np.arange(T)[None, :, None] < lens[:, None, None]
    aud = rng.uniform(-1, 1, (B, T, A)).astype(np.float32)
    vis = rng.uniform(-1, 1, (B, T, Vd)).astype(np.float32)
    aud_m = np.broadcast_to(step, aud.shape).astype(np.float32).copy()
    vis_m = np.broadcast_to(step, vis.shape).astype(np.float32).copy()
    # sprinkle in-sequence zeros (normalize_data masks exact zeros feature-wise)
    hole_a = rng.random(aud.shape) < 0.05
    hole_v = rng.random(vis.shape) < 0.05
    aud_m[hole_a] = 0.0
    vis_m[hole_v] = 0.0
    aud[aud_m == 0] = -10.0
    vis[vis_m == 0] = -10.0
    latents = (We[ids] * ww[ids][:, :, None]).mean(1).astype(np.float32)
    latents += 0.05 * rng.standard_normal(latents.shape).astype(np.float32)
    heads = {}
    for mod, D in head_dims(d, A, Vd, unimodal).items():
        heads[mod] = tuple(x.astype(np.float32) for x in (
            rng.uniform(-1, 1, (D, d)) / np.sqrt(d), rng.uniform(-1, 1, D) / np.sqrt(d),
            rng.uniform(-1, 1, (D, d)) / np.sqrt(d), rng.uniform(-1, 1, D) / np.sqrt(d)))
    out = dict(d=d, A=A, Vd=Vd, We=We, ids=ids, text=text, text_m=text_m, text_w=ww[ids],
               aud=aud, vis=vis, aud_m=aud_m, vis_m=vis_m, latents=latents, heads=heads)
    if norm is not None:
        out['norm_params'] = ((1.0 + 0.1 * rng.standard_normal(d)).astype(np.float32),
                              (0.1 * rng.standard_normal(d)).astype(np.float32))
    return out


def load_heads(model, heads, norm_params=None):
    """Copy the seeded head parameters into an ``AudioVisualGeneratorMultimodal``."""
    import torch
    with torch.no_grad():
        for mod, (Wm, bm, Ws, bs) in heads.items():
            model.embed2out[mod]['mu'].weight.copy_(torch.tensor(Wm))
            model.embed2out[mod]['mu'].bias.copy_(torch.tensor(bm))
            model.embed2out[mod]['log_sigma'].weight.copy_(torch.tensor(Ws))
            model.embed2out[mod]['log_sigma'].bias.copy_(torch.tensor(bs))
        if norm_params is not None and model.norm is not None:
            model.norm.weight.copy_(torch.tensor(norm_params[0]))
            model.norm.bias.copy_(torch.tensor(norm_params[1]))


# optimize_latents (reference simplesif.py:49-162) end-to-end cases: fixed batch order.
OPT_CASES = {
    'sgd_train': dict(inputs=dict(B=12, T=5, d=24, A=6, Vd=5, V=40, seed=41, norm='layer_norm', unimodal=False,
                                  args={}),
                      args={'dataset': 'mosi', 'unimodal': False, 'freeze_weights': False, 'optimizer': 'sgd',
                            'word_loss_weight': 0.2},
                      train=True, batch=4, epochs=4, lr=0.01),
    'adam_infer': dict(inputs=dict(B=10, T=4, d=24, A=5, Vd=4, V=30, seed=42, norm=None, unimodal=True, args={}),
                       args={'dataset': 'mosi', 'unimodal': True, 'freeze_weights': True, 'optimizer': 'adam'},
                       train=False, batch=5, epochs=3, lr=0.01),
    # BatchNorm1d in training mode (batch statistics, running buffers updated every step): half of the
    # reference's grid (make_configs.py:30)
    'sgd_train_bn': dict(inputs=dict(B=12, T=5, d=24, A=6, Vd=5, V=40, seed=43, norm='batch_norm', unimodal=False,
                                     args={}),
                         args={'dataset': 'mosi', 'unimodal': False, 'freeze_weights': False, 'optimizer': 'sgd',
                               'word_loss_weight': 0.2},
                         train=True, batch=4, epochs=4, lr=0.01),
    'adam_infer_bn': dict(inputs=dict(B=12, T=4, d=24, A=5, Vd=4, V=30, seed=44, norm='batch_norm', unimodal=False,
                                      args={}),
                          args={'dataset': 'mosi', 'unimodal': False, 'freeze_weights': False, 'optimizer': 'adam',
                                'word_loss_weight': 0.1},
                          train=False, batch=6, epochs=3, lr=0.01),
}


def raw_split(n=7, T=6, A=9, Vd=5, seed=51):
    """A raw (un-normalised) split like the reference's h5 arrays: exact zeros mark padding,
    audio feature 3 is constant (normalize_data drops it)."""
    rng = np.random.default_rng(seed)
    cov = rng.normal(size=(n, T, A)) * 3 + 1
    fac = rng.normal(size=(n, T, Vd)) * 2 - 1
    lens = rng.integers(2, T + 1, size=n)
    pad = np.arange(T)[None, :, None] >= lens[:, None, None]
    cov[:, :, 3] = 0.0
    cov[np.broadcast_to(pad, cov.shape)] = 0.0
    fac[np.broadcast_to(pad, fac.shape)] = 0.0
    return {'covarep': cov, 'facet': fac}


CLOSED_FORM_CASES = {
    # estimate_embedding_overall_gpu2 (reference sif2.py:164-208) on one split of mmb_inputs
    'mmb2_small': dict(B=9, T=6, d=24, A=6, Vd=5, V=40, seed=61, norm=None, unimodal=False, args={}),
    'mmb2_mosi': dict(B=70, T=20, d=300, A=76, Vd=49, V=300, seed=62, norm=None, unimodal=False, args={}),
}
CLOSED_FORM_KEYS = ['audio', 'visual', 'audiovisual', 'textaudio', 'textvisual', 'textaudiovisual']


def closed_form_data(c):
    """The concatenated modalities the reference's call site builds (simplesif.py:826-846)."""
    cat = lambda *xs: np.concatenate(xs, axis=-1)
    return {'audio': c['aud'], 'visual': c['vis'], 'audiovisual': cat(c['aud'], c['vis']),
            'textaudio': cat(c['text'], c['aud']), 'textvisual': cat(c['text'], c['vis']),
            'textaudiovisual': cat(c['text'], c['aud'], c['vis'])}


# Downstream parity (north_star: "MOSI/POM MAE and correlation unchanged to the 3rd decimal"): the
# whole script path -- SIF initialisation per split, latent optimisation, regressor -- on small
# labelled synthetic splits, run through the reference (make_golden.golden_downstream) and through
# this repo (tests/test_mmb_gpu.py) from the same seed.
_DS_ARGS = {'unimodal': False, 'freeze_weights': False, 'word_sim_metric': 'angular', 'batch_size': 64,
            'n_runs': 1, 'semi_sup_idxes': None, 'config_name': 'golden', 'config_num': 0, 'lr_decay': 0.5,
            'early_stopping': False, 'time_test': False, 'seq_len': 20, 'sentiment_hidden_size': 100,
            'word_loss_weight': 0.002, 'likelihood_weight': 0.001, 'pos_embed_dim': 2}
DOWNSTREAM_CASES = {
    # the e2e branch (reference simplesif.py:625-806; every generated config has e2e=True)
    'mosi_e2e': dict(n=(160, 48, 80), L=12, T_a=None, V=400, d=300, A=10, Vd=8, n_out=1, seed=71,
                     args=dict(_DS_ARGS, dataset='mosi', e2e=True, optimizer='sgd', norm='layer_norm', lr=1e-3,
                               sentiment_lr=0.1, n_epochs=12, n_sentiment_epochs=41)),
    # the two-stage branch (simplesif.py:541-624) on the POM layout: unaligned ids for the word term,
    # aligned word vectors for the Gaussian terms (MMDataExtra), several traits -> pom_loss
    'pom_two_stage': dict(n=(130, 40, 70), L=16, T_a=24, V=300, d=300, A=9, Vd=7, n_out=3, seed=72,
                          args=dict(_DS_ARGS, dataset='pom', e2e=False, optimizer='adam', norm=None, lr=1e-3,
                                    sentiment_lr=0.1, n_epochs=11, n_sentiment_epochs=31)),
}


def downstream_inputs(n, L, T_a, V, d, A, Vd, n_out, seed, args):
    """Three labelled splits in the layout the reference's loaders + normalize_data hand to main():
    ids (N, L) right-padded with 0, covarep / facet in [-1, 1] with padded steps -10 and int masks
    (utils.py:171-189), labels that depend on the words (and the first audio feature on the labels)."""
    rng = np.random.default_rng(seed)
    We = table(V, d, seed + 100)
    U = rng.standard_normal((d, n_out))
    splits, masks, weights = [], [], None
    for k, nk in enumerate(n):
        ids, p = zipf_ids(rng, nk, L, V, lo_len=3)
        if weights is None:
            weights = sif_weights(p)
        lens = (ids != 0).sum(1)
        T = L if T_a is None else T_a
        if T_a is None:
            step_len = lens
        else:   # aligned stream: every word held for 1..2 frames, cut to T_a
            rep = rng.integers(1, 3, size=ids.shape)
            ids_a = np.zeros((nk, T_a), dtype=np.int64)
            for i in range(nk):
                row = np.repeat(ids[i], rep[i])
                row = row[row != 0][:T_a]
                ids_a[i, :row.size] = row
            step_len = (ids_a != 0).sum(1)
        step = np.arange(T)[None, :, None] < step_len[:, None, None]
        avg = (We[ids] * weights[ids][:, :, None]).sum(1) / L           # labels: a noisy function of the
        z = (avg - avg.mean(0)) @ U                                       # (centred) weighted average
        y = 3.0 * np.tanh(0.7 * z / z.std(0)) + 0.3 * rng.standard_normal((nk, n_out))
        cov = rng.uniform(-1, 1, (nk, T, A))
        fac = rng.uniform(-1, 1, (nk, T, Vd))
        cov[:, :, 0] = np.clip(cov[:, :, 0] * 0.5 + y[:, :1] / 6.0, -1, 1)
        cov_m = np.broadcast_to(step, cov.shape).astype(np.int64).copy()
        fac_m = np.broadcast_to(step, fac.shape).astype(np.int64).copy()
        cov[cov_m == 0] = -10.0
        fac[fac_m == 0] = -10.0
        s = {'covarep': cov, 'facet': fac, 'label': (y[:, 0] if n_out == 1 else y).astype(np.float32)}
        if T_a is None:
            s['text'] = ids
        else:
            s['text_id'] = ids
            s['text'] = We[ids_a].astype(np.float32)          # aligned word vectors (zero rows = padding)
        splits.append(s)
        masks.append({'covarep': cov_m, 'facet': fac_m})
    return We, weights, splits, masks


# sentiment_model.train_sentiment_for_latents alone (pure torch, runs on the CPU on both sides): fixed
# random latents, labels a noisy function of them.
SENTIMENT_CASES = {
    'mosi': dict(dataset='mosi', n_out=1, early_stopping=False, seed=81),
    'pom': dict(dataset='pom', n_out=3, early_stopping=False, seed=82),
    'mosi_early_stopping': dict(dataset='mosi', n_out=1, early_stopping=True, seed=83),
    # (N, 1) label layout: the reference's squeezed (B,) prediction broadcasts against (B, 1) labels
    # to a (B, B) L1 matrix (sentiment_model.py:90-103) -- reproduced, see sentiment_model._l1
    'mosi_column_labels': dict(dataset='mosi', n_out=1, early_stopping=False, seed=84, column_labels=True),
}


def sentiment_inputs(dataset, n_out, early_stopping, seed, d=30, sizes=(150, 45, 70), column_labels=False):
    rng = np.random.default_rng(seed)
    lat = [rng.standard_normal((n, d)).astype(np.float32) for n in sizes]
    U = rng.standard_normal((d, n_out))
    labs = [(np.tanh(x @ U) * 3 + 0.3 * rng.standard_normal((len(x), n_out))).astype(np.float32) for x in lat]
    if n_out == 1 and not column_labels:
        labs = [y[:, 0] for y in labs]
    args = {'dataset': dataset, 'sentiment_hidden_size': 20, 'n_sentiment_epochs': 95, 'sentiment_lr': 0.1,
            'early_stopping': early_stopping, 'lr_decay': 0.5}
    return args, lat, labs
