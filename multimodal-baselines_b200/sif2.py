"""Drop-in for the live part of the reference's ``sif2.py``: the closed-form
multimodal latent estimate the reference times under ``--time_test`` (simplesif.py:808-880).

``calc_weights`` (reference sif2.py:103-114) and ``estimate_embedding_overall_gpu2`` (164-208)
keep their names, argument order and return values; the estimate runs in libmmb_b200.so as one
pass over the base tensors for the time sums, one split-K product with the head weights and one
normalisation kernel.  The rest of the reference's sif2.py is a stale second driver that cannot
run (undefined names, old signatures -- SURVEY.md §2 #13) and is not reproduced.
"""
import ctypes as C

import torch

import _native as nv
import mmb_ops
from _native import lib
from losses import CatSegments

KEYS = ['audio', 'visual', 'audiovisual', 'textaudio', 'textvisual', 'textaudiovisual']


def calc_weights(data, b_mean, b_log_sigma, mask):
    """reference sif2.py:103-114 (``mask`` is accepted and, as in the reference, unused)."""
    b_mean = b_mean.reshape((1, 1, -1))
    b_log_sigma = b_log_sigma.reshape((1, 1, -1))
    inv = torch.exp(-2 * b_log_sigma)
    q_mean = (data - b_mean) * inv
    q_sigma = (data - b_mean) ** 2 * inv - 1.
    return q_mean, q_sigma


def _parts(x):
    return list(x.parts) if isinstance(x, CatSegments) else [x]


def estimate_embedding_overall_gpu2(data, masks, networks, sentence_weights, embeddings, keys=None):
    """reference sif2.py:164-208.

    data[k]: (N, T, D_k) tensor or ``CatSegments`` of base tensors; networks[k] = (mu, log_sigma)
    ``nn.Linear`` pair of head k; sentence_weights (N, L); embeddings (N, L, d) word vectors.
    Returns the (N, d) unit-norm latent estimates.  ``masks`` is unused, as in the reference.
    """
    keys = list(keys) if keys is not None else [k for k in KEYS if k in data]
    dev = nv.require_cuda()
    f32 = mmb_ops._f32
    sent_w = f32(sentence_weights)
    emb = f32(embeddings)
    N, L = sent_w.shape
    d = emb.shape[-1]
    segs, Fs, n_seg, Ds = [], [], [], []
    T = None
    for k in keys:
        parts = [f32(p) for p in _parts(data[k])]
        n_seg.append(len(parts))
        D = 0
        for p in parts:
            if p.dim() != 3 or p.shape[0] != N:
                raise ValueError('data[%s] must be (N, T, features)' % k)
            T = p.shape[1] if T is None else T
            if p.shape[1] != T:
                raise ValueError('all modalities must share the time axis')
            segs.append(p)
            Fs.append(p.shape[2])
            D += p.shape[2]
        Ds.append(D)
    b_mu = [f32(networks[k][0].bias.detach()) for k in keys]
    b_ls = [f32(networks[k][1].bias.detach()) for k in keys]
    W_mu = [f32(networks[k][0].weight.detach()) for k in keys]
    W_ls = [f32(networks[k][1].weight.detach()) for k in keys]
    for k, D, w in zip(keys, Ds, W_mu):
        if w.shape != (D, d):
            raise RuntimeError('head %s: weight %s does not match data width %d' % (k, tuple(w.shape), D))
    S1 = [torch.empty((N, D), dtype=torch.float32, device=dev) for D in Ds]
    S2 = [torch.empty((N, D), dtype=torch.float32, device=dev) for D in Ds]
    tw_part = torch.empty((len(keys), N), dtype=torch.float32, device=dev)
    P, I = mmb_ops._ptr_array, mmb_ops._int_array
    nv.check(lib.mmb_closed_form_stats(N, T, len(keys), I(n_seg), P(segs), I(Fs), P(b_mu), P(b_ls), P(S1), P(S2),
                                       nv.ptr(tw_part), nv.stream_ptr()))
    # prod = sum_k S1_k W_mu_k + S2_k W_ls_k : the head input-gradient product with gout = (S1, S2)
    gouts, Ws, Dh = [], [], []
    for i in range(len(keys)):
        gouts += [S1[i], S2[i]]
        Ws += [W_mu[i], W_ls[i]]
        Dh += [Ds[i], Ds[i]]
    prod = torch.empty((N, d), dtype=torch.float32, device=dev)
    Dc = I(Dh)
    nbytes = lib.mmb_heads_backward_workspace_bytes(N, d, len(Ws), Dc)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    nv.check(lib.mmb_heads_backward(nv.ptr(emb), N, d, len(Ws), P(Ws), Dc, P(gouts), nv.ptr(prod), None, None,
                                    nv.ptr(ws), nbytes, nv.stream_ptr()))
    out = torch.empty((N, d), dtype=torch.float32, device=dev)
    nv.check(lib.mmb_closed_form_finish(N, L, d, nv.ptr(sent_w), nv.ptr(emb), nv.ptr(prod), nv.ptr(tw_part),
                                        len(keys), nv.ptr(out), nv.stream_ptr()))
    return out
