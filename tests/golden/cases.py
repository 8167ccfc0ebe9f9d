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
