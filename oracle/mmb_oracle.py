"""CPU oracle for the MMB half of the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/sif_oracle.py`` for the import rule).

NumPy float64 restatement of the reference's PyTorch forward passes, each function citing
the reference ``file:line`` it follows, plus the analytic gradients (SURVEY.md Appendix
A.3-A.6) so the CUDA backward kernels can be checked without autograd.  Pinned against
outputs (values and autograd gradients) of the reference's own ``losses.py`` / ``models.py``
run in the build container: ``tests/golden/make_golden.py`` -> ``tests/golden/mmb_*.npz``.
Third-party arithmetic: PyTorch (2.11.0 here) ``nn.CosineSimilarity`` (per-vector norm
clamp at eps=1e-8), ``acos``, ``nn.Linear``, ``nn.LayerNorm`` (eps=1e-5, biased variance),
``nn.BatchNorm1d`` in training mode (batch statistics, biased variance, eps=1e-5).
"""
from __future__ import annotations

import numpy as np

LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)


# --------------------------------------------------------------------------- A7
def get_normal_log_prob(mu, sigma, values, mask):
    """losses.py:13-34.

    ``term1 = log(1/sqrt(2 pi sigma^2))`` (26), ``term2 = (x-mu)^2/(2 sigma^2)`` (28-29),
    ``(term1-term2)*mask`` summed over the last two axes (31-33).
    mu, sigma: (B,1,D) or (B,D); values, mask: (B,T,D).  Returns (B,).
    """
    mu = np.asarray(mu, dtype=np.float64)
    sigma = np.asarray(sigma, dtype=np.float64)
    if mu.ndim == 2:
        mu = mu[:, None, :]
        sigma = sigma[:, None, :]
    values = np.asarray(values, dtype=np.float64)
    mask = np.asarray(mask, dtype=np.float64)
    sig_sq = sigma ** 2
    term1 = np.log(1.0 / np.sqrt(2.0 * np.pi * sig_sq))
    term2 = (values - mu) ** 2 / (2.0 * sig_sq)
    return ((term1 - term2) * mask).sum(-1).sum(-1)


def normal_log_prob_grads(mu, log_sigma, values, mask):
    """Appendix A.4: gradients of A7 w.r.t. ``mu`` and ``s = log sigma`` (sigma = exp s,
    models.py:199).  ``d/dmu = sum_t m (x-mu)/sigma^2``; ``d/ds = sum_t m (-1 + (x-mu)^2/sigma^2)``.
    Returns (lp (B,), dmu (B,D), ds (B,D))."""
    mu = np.asarray(mu, dtype=np.float64)
    s = np.asarray(log_sigma, dtype=np.float64)
    x = np.asarray(values, dtype=np.float64)
    m = np.asarray(mask, dtype=np.float64)
    inv_var = np.exp(-2.0 * s)[:, None, :]
    diff = x - mu[:, None, :]
    lp = (m * (-(LOG_SQRT_2PI + s[:, None, :]) - 0.5 * diff * diff * inv_var)).sum((1, 2))
    dmu = (m * diff * inv_var).sum(1)
    ds = (m * (-1.0 + diff * diff * inv_var)).sum(1)
    return lp, dmu, ds


def gauss_moments(values, mask):
    """Masked moments over time: ``S0 = sum_t m``, ``mean = sum_t m x / S0``,
    ``M2 = sum_t m (x - mean)^2``, stacked (B, 3, D); ``S0 = 0`` gives mean = M2 = 0.  The identity the CUDA
    path (mmb_gauss_moments / mmb_gauss_ll_stats) rests on: ``sum_t m (x - mu)^2 = M2 + S0 (mean - mu)^2``."""
    x = np.asarray(values, dtype=np.float64)
    m = np.asarray(mask, dtype=np.float64)
    s0 = m.sum(1)
    with np.errstate(invalid='ignore', divide='ignore'):
        mean = np.where(s0 != 0, (m * x).sum(1) / s0, 0.0)
    m2 = (m * (x - mean[:, None, :]) ** 2).sum(1)
    return np.stack([s0, mean, m2], axis=1)


def normal_log_prob_from_moments(mu, log_sigma, stats):
    """A7 and its gradients (as ``normal_log_prob_grads``) evaluated from ``gauss_moments`` alone."""
    mu = np.asarray(mu, dtype=np.float64)
    s = np.asarray(log_sigma, dtype=np.float64)
    s0, mean, m2 = (np.asarray(stats, dtype=np.float64)[:, k, :] for k in range(3))
    inv_var = np.exp(-2.0 * s)
    q = m2 + s0 * (mean - mu) ** 2
    lp = (-(LOG_SQRT_2PI + s) * s0 - 0.5 * q * inv_var).sum(-1)
    return lp, s0 * (mean - mu) * inv_var, -s0 + q * inv_var


# --------------------------------------------------------------------------- A8
def _cos_rows(a, b, eps=1e-8):
    """torch ``nn.CosineSimilarity(dim=-1)``: each vector divided by ``max(norm, eps)``."""
    na = np.maximum(np.linalg.norm(a, axis=-1, keepdims=True), eps)
    nb = np.maximum(np.linalg.norm(b, axis=-1, keepdims=True), eps)
    return ((a / na) * (b / nb)).sum(-1)


def get_word_log_prob_angular2(latents, word_embeddings, word_weights, sent_embeddings, mask, a):
    """losses.py:68-95.

    ``Z = sum_v (1 - acos(cos(e, We_v))/pi)`` (74-76); ``alpha = 1/(Z a + 1)`` (78);
    per token ``log(alpha w_t + (1-alpha)(1 - acos(cos(We_{x_t}, e))/pi)/Z)`` (80-87),
    masked by ``mask[:,:,0]`` (90) and summed over tokens (92).
    latents (B,d), word_embeddings (V,d), word_weights (B,L), sent_embeddings (B,L,d),
    mask (B,L,d) or (B,L).  Returns (B,).
    """
    e = np.asarray(latents, dtype=np.float64)
    W = np.asarray(word_embeddings, dtype=np.float64)
    ww = np.asarray(word_weights, dtype=np.float64)
    S = np.asarray(sent_embeddings, dtype=np.float64)
    mask = np.asarray(mask, dtype=np.float64)
    m = mask[:, :, 0] if mask.ndim == 3 else mask
    cos_v = _cos_rows(e[:, None, :], W[None, :, :])
    Z = (1.0 - np.arccos(cos_v) / np.pi).sum(-1, keepdims=True)
    alpha = 1.0 / (Z * a + 1.0)
    unigram = alpha * ww
    score = 1.0 - np.arccos(_cos_rows(S, e[:, None, :])) / np.pi
    context = (1.0 - alpha) * score / Z
    return (np.log(unigram + context) * m).sum(-1)


def word_log_prob_grad(latents, word_embeddings, word_weights, sent_embeddings, mask, a, eps=1e-8):
    """Appendix A.5: d lp / d latents, (B,d).  Assumes no norm hits the 1e-8 clamp except
    all-zero word rows (whose cosine and gradient contribution are exactly 0)."""
    e = np.asarray(latents, dtype=np.float64)
    W = np.asarray(word_embeddings, dtype=np.float64)
    ww = np.asarray(word_weights, dtype=np.float64)
    S = np.asarray(sent_embeddings, dtype=np.float64)
    mask = np.asarray(mask, dtype=np.float64)
    m = mask[:, :, 0] if mask.ndim == 3 else mask
    ne = np.maximum(np.linalg.norm(e, axis=-1, keepdims=True), eps)
    eh = e / ne
    Wh = W / np.maximum(np.linalg.norm(W, axis=-1, keepdims=True), eps)
    Sh = S / np.maximum(np.linalg.norm(S, axis=-1, keepdims=True), eps)
    C = eh @ Wh.T                                            # (B,V)
    Z = (1.0 - np.arccos(C) / np.pi).sum(-1, keepdims=True)  # (B,1)
    alpha = 1.0 / (a * Z + 1.0)
    ct = np.einsum('bld,bd->bl', Sh, eh)
    st = 1.0 - np.arccos(ct) / np.pi
    p = alpha * ww + (1.0 - alpha) * st / Z
    dlp_dp = m / p
    dalpha_dZ = -a * alpha * alpha
    dp_dZ = dalpha_dZ * (ww - st / Z) - (1.0 - alpha) * st / (Z * Z)
    dlp_dZ = (dlp_dp * dp_dZ).sum(-1, keepdims=True)         # (B,1)
    dlp_ds = dlp_dp * (1.0 - alpha) / Z                      # (B,L)
    with np.errstate(divide='ignore', invalid='ignore'):
        R = dlp_dZ / (np.pi * np.sqrt(1.0 - C * C))          # (B,V)
        r = dlp_ds / (np.pi * np.sqrt(1.0 - ct * ct))        # (B,L)
    g = (R @ Wh - (R * C).sum(-1, keepdims=True) * eh) / ne
    g += (np.einsum('bl,bld->bd', r, Sh) - (r * ct).sum(-1, keepdims=True) * eh) / ne
    return g


def word_log_prob_grad_from_ids(latents, word_embeddings, word_weights, ids, mask, a, eps=1e-8):
    """The algebra of mmb_word_ll_ids restated: when the token vectors are rows of the
    table (``sent = W[ids]``), the token cosines are entries of the (B, V) cosine matrix ``C`` of the partition
    term and the token part of the gradient is a scatter into the (B, V) coefficient matrix ``M`` of the one
    product ``M @ W``:  ``M[b,v] = DZ_b h(C[b,v]) iw_v + sum_{t: id_t = v} r_t iw_v``,
    ``grad_b = (M_b @ W - (DZ_b HC_b + RC_b) e_hat_b) / ||e_b||``.  Returns (lp (B,), grad (B,d))."""
    e = np.asarray(latents, dtype=np.float64)
    W = np.asarray(word_embeddings, dtype=np.float64)
    ww = np.asarray(word_weights, dtype=np.float64)
    ids = np.asarray(ids)
    mask = np.asarray(mask, dtype=np.float64)
    m = mask[:, :, 0] if mask.ndim == 3 else mask
    ne = np.maximum(np.linalg.norm(e, axis=-1, keepdims=True), eps)
    eh = e / ne
    iw = 1.0 / np.maximum(np.linalg.norm(W, axis=-1), eps)
    C = (e @ W.T) / ne * iw[None, :]
    S = 1.0 - np.arccos(C) / np.pi
    with np.errstate(divide='ignore', invalid='ignore'):
        H = 1.0 / (np.pi * np.sqrt(1.0 - C * C))
    Z = S.sum(-1, keepdims=True)
    HC = (H * C).sum(-1)
    alpha = 1.0 / (a * Z + 1.0)
    rows = np.arange(e.shape[0])[:, None]
    ct, st = C[rows, ids], S[rows, ids]
    p = alpha * ww + (1.0 - alpha) * st / Z
    lp = (np.log(p) * m).sum(-1)
    dlp_dp = m / p
    dp_dZ = (-a * alpha * alpha) * (ww - st / Z) - (1.0 - alpha) * st / (Z * Z)
    DZ = (dlp_dp * dp_dZ).sum(-1)
    with np.errstate(divide='ignore', invalid='ignore'):
        rr = dlp_dp * (1.0 - alpha) / Z / (np.pi * np.sqrt(1.0 - ct * ct))
    RC = (rr * ct).sum(-1)
    M = DZ[:, None] * H * iw[None, :]
    np.add.at(M, (np.broadcast_to(rows, ids.shape), ids), rr * iw[ids])
    grad = (M @ W - (DZ * HC + RC)[:, None] * eh) / ne
    return lp, grad


# --------------------------------------------------------------------------- A6
MMB1_MODALITIES = ('audio', 'visual')
MMB2_MODALITIES = ('audio', 'visual', 'audiovisual', 'textaudio', 'textvisual', 'textaudiovisual')


def modality_dims(embedding_dim, audio_dim, visual_dim, unimodal=False):
    """models.py:115-159: output width of every head, in ModuleDict order."""
    if unimodal:
        return {'audio': audio_dim, 'visual': visual_dim}
    return {
        'audio': audio_dim,
        'visual': visual_dim,
        'audiovisual': audio_dim + visual_dim,
        'textaudio': embedding_dim + audio_dim,
        'textvisual': embedding_dim + visual_dim,
        'textaudiovisual': embedding_dim + audio_dim + visual_dim,
    }


def apply_norm(e, norm, gamma=None, beta=None, eps=1e-5):
    """models.py:161-168, 188-191: ``None`` | ``'layer_norm'`` (nn.LayerNorm(d)) |
    ``'batch_norm'`` (nn.BatchNorm1d(d), training mode -> batch statistics, biased var)."""
    e = np.asarray(e, dtype=np.float64)
    if norm is None:
        return e
    if norm == 'layer_norm':
        mean = e.mean(-1, keepdims=True)
        var = e.var(-1, keepdims=True)
    elif norm == 'batch_norm':
        mean = e.mean(0, keepdims=True)
        var = e.var(0, keepdims=True)
    else:
        raise NotImplementedError
    z = (e - mean) / np.sqrt(var + eps)
    if gamma is not None:
        z = z * np.asarray(gamma, dtype=np.float64) + np.asarray(beta, dtype=np.float64)
    return z


def heads_forward(e, params, norm=None, norm_params=None):
    """models.py:187-202: ``mu = z W_mu^T + b_mu``; ``sigma = exp(z W_ls^T + b_ls)``.
    ``params[mod] = (W_mu, b_mu, W_ls, b_ls)``; returns ``{mod: {'mu','sigma'}}``."""
    g, b = norm_params if norm_params is not None else (None, None)
    z = apply_norm(e, norm, g, b)
    out = {}
    for mod, (Wm, bm, Ws, bs) in params.items():
        out[mod] = {
            'mu': z @ np.asarray(Wm, np.float64).T + np.asarray(bm, np.float64),
            'sigma': np.exp(z @ np.asarray(Ws, np.float64).T + np.asarray(bs, np.float64)),
        }
    return out


# --------------------------------------------------------------------------- A9
def get_log_prob_matrix(args, latents, out, data, masks, word_log_prob_fn):
    """losses.py:216-274: word term first (236), then one A7 per modality (251-256);
    ``'word_loss_weight' in args`` -> ``sum_m lp_m (1-w)/M + w lp_word`` (267-270) else
    plain sum (272).  The reference's inf check + ``sys.exit()`` (258-264) is reported
    here by raising ``FloatingPointError``."""
    word_lp = word_log_prob_fn(latents, data['text_weights'], data['text'], masks['text'])
    log_probs = {}
    for modality, d in out.items():
        log_probs[modality] = get_normal_log_prob(d['mu'], d['sigma'], data[modality], masks[modality])
    for m, lp in log_probs.items():
        if np.abs(lp.min()) == np.inf:
            raise FloatingPointError(m + ' inf')
    total = sum(log_probs.values())
    if 'word_loss_weight' in args:
        w = args['word_loss_weight']
        return total * ((1.0 - w) / len(log_probs)) + w * word_lp
    return total + word_lp


def concat_modalities(text_gauss, aud, vis, unimodal=False):
    """simplesif.py:94-124: the per-step ``torch.cat`` views of the three base tensors."""
    d = {'audio': aud, 'visual': vis}
    if not unimodal:
        d['audiovisual'] = np.concatenate([aud, vis], -1)
        d['textaudio'] = np.concatenate([text_gauss, aud], -1)
        d['textvisual'] = np.concatenate([text_gauss, vis], -1)
        d['textaudiovisual'] = np.concatenate([text_gauss, aud, vis], -1)
    return d


# --------------------------------------------------------------------------- closed form (N1)
def calc_weights(data, b_mean, b_log_sigma, mask=None):
    """reference sif2.py:103-114: q_mean = (x - b_mu) / exp(2 b_ls), q_sigma = (x - b_mu)^2 /
    exp(2 b_ls) - 1 (the mask argument is unused there too)."""
    data = np.asarray(data, dtype=np.float64)
    bm = np.asarray(b_mean, dtype=np.float64).reshape(1, 1, -1)
    bl = np.asarray(b_log_sigma, dtype=np.float64).reshape(1, 1, -1)
    return (data - bm) / np.exp(2 * bl), (data - bm) ** 2 / np.exp(2 * bl) - 1.0


def estimate_embedding_overall(data, heads, sentence_weights, embeddings, keys):
    """reference sif2.py:164-208 (estimate_embedding_overall_gpu2) in float64.
    data[k] (N, T, D_k); heads[k] = (W_mu, b_mu, W_ls, b_ls); sentence_weights (N, L);
    embeddings (N, L, d).  Returns (N, d) unit rows."""
    sw = np.asarray(sentence_weights, dtype=np.float64)
    E = np.asarray(embeddings, dtype=np.float64)
    qm, qs = {}, {}
    for k in keys:
        Wm, bm, Ws, bs = heads[k]
        qm[k], qs[k] = calc_weights(data[k], bm, bs)
    tw = sw.sum(-1) + sum(q.sum(-1).sum(-1) for q in qm.values()) + sum(q.sum(-1).sum(-1) for q in qs.values())
    cs = np.einsum('nl,nld->nd', sw, E)
    for k in keys:
        Wm, bm, Ws, bs = heads[k]
        cs = cs + qm[k].sum(1) @ np.asarray(Wm, dtype=np.float64) + qs[k].sum(1) @ np.asarray(Ws, dtype=np.float64)
    cs = cs / tw[:, None]
    return cs / np.linalg.norm(cs, axis=1, keepdims=True)
