"""CPU oracle for the SIF half of the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``multimodal-baselines_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, and there only as the checker / the timed CPU baseline.

It is a NumPy restatement of the reference's algorithm, every function citing the
reference ``file:line`` it follows.  The one piece of arithmetic that is not in the
reference tree is scikit-learn's ``TruncatedSVD`` (called at sif_functions.py:65-66 with
``n_components=npc, n_iter=7, random_state=0``).  The reference pins no version; the
version that defines the oracle is the one in this image, scikit-learn 1.9.0
(``sklearn/decomposition/_truncated_svd.py`` -> ``sklearn/utils/extmath.py``
``_randomized_svd`` / ``_randomized_range_finder`` / ``svd_flip``).  ``compute_pc`` below
calls it directly; ``compute_pc_restated`` restates its published algorithm in NumPy/SciPy
so the two can be checked against each other, and ``compute_pc_from_gram`` is the closed
form in the 300x300 Gram that the CUDA solver implements.

Parity pinning: the reference has no tests and no golden vectors, so this
oracle is pinned against outputs of the reference itself, generated in the build container
by ``tests/golden/make_golden.py`` (which imports /root/reference unmodified) and committed
as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "seq2weight", "seq2weight_loop", "get_weighted_average", "get_weighted_average_loop",
    "compute_pc", "compute_pc_restated", "compute_pc_from_gram", "remove_pc",
    "remove_pc_with", "SIF_embedding", "get_sentence_embeddings",
    "get_sentence_embeddings_loop", "Params", "get_word_weights", "start_block",
]


class Params(object):
    """sif_functions.py:17-26 -- plain attribute bag; ``rmpc`` is set by the caller."""

    def __init__(self):
        self.LW = 1e-5
        self.LC = 1e-5
        self.eta = 0.05


# --------------------------------------------------------------------------- A1
def seq2weight(seq, mask, weight4ind):
    """sif_functions.py:8-15, vectorised.

    ``w[i,j] = float32(weight4ind[seq[i,j]])`` where ``mask[i,j] > 0 and seq[i,j] >= 0``,
    else 0.  The float64 -> float32 cast happens on assignment into the float32 array
    (line 9 allocates float32, line 13 stores), i.e. round-to-nearest-even.
    """
    seq = np.asarray(seq)
    weight4ind = np.asarray(weight4ind)
    take = (np.asarray(mask) > 0) & (seq >= 0)
    out = np.zeros(seq.shape, dtype=np.float32)
    # weight4ind[seq] raises IndexError for seq >= len(weight4ind), as the reference does.
    out[take] = weight4ind[seq[take]].astype(np.float32)
    return out


def seq2weight_loop(seq, mask, weight4ind):
    """sif_functions.py:8-15 with the reference's own per-element Python double loop.

    Kept because that loop is ~60 % of the reference's CPU time at scale;
    ``bench.py``'s CPU baseline times this variant so the baseline is the reference's cost.
    """
    weight = np.zeros(seq.shape).astype('float32')
    for i in range(seq.shape[0]):
        for j in range(seq.shape[1]):
            if mask[i, j] > 0 and seq[i, j] >= 0:
                weight[i, j] = weight4ind[seq[i, j]]
    return np.asarray(weight, dtype='float32')


# --------------------------------------------------------------------------- A2
def get_weighted_average_loop(We, x, w):
    """sif_functions.py:28-56, the reference's per-utterance loop (line 54-55).

    ``emb`` is float64 (line 37: ``np.zeros`` default dtype); each row is
    ``w[i,:].dot(We[x[i,:],:]) / count_nonzero(w[i,:])`` evaluated in the result type of
    the dot (float32 when both ``w`` and ``We`` are float32).
    """
    n_samples = x.shape[0]
    emb = np.zeros((n_samples, We.shape[1]))
    for i in range(n_samples):
        emb[i, :] = w[i, :].dot(We[x[i, :], :]) / np.count_nonzero(w[i, :])
    return emb


def get_weighted_average(We, x, w, accumulate=np.float64):
    """sif_functions.py:28-56, blocked and accumulated in ``accumulate`` precision.

    The float64 accumulation is the *tolerance reference* (embedding parity is stated
    relative to the exact sum, 1e-5); ``get_weighted_average_loop`` is the bit-faithful
    restatement.  Divisor quirk kept: ``count_nonzero`` of the float32 weights over the
    whole padded row (pad id 0 has a non-zero weight, so it counts).  A row whose weights
    are all zero divides by zero -> NaN, as in the reference.
    """
    We = np.asarray(We)
    x = np.asarray(x)
    w = np.asarray(w)
    n, L = x.shape
    d = We.shape[1]
    emb = np.zeros((n, d), dtype=np.float64)
    block = max(1, (1 << 24) // max(1, L * d))
    for s in range(0, n, block):
        xs = x[s:s + block]
        ws = w[s:s + block].astype(accumulate)
        rows = We[xs].astype(accumulate)                 # (b, L, d); negative ids wrap
        acc = np.einsum('bl,bld->bd', ws, rows)
        cnt = np.count_nonzero(w[s:s + block], axis=1).astype(accumulate)
        with np.errstate(divide='ignore', invalid='ignore'):
            emb[s:s + block] = acc / cnt[:, None]
    return emb


# --------------------------------------------------------------------------- A3
def compute_pc(X, npc=1):
    """sif_functions.py:58-67 -- sklearn ``TruncatedSVD(npc, n_iter=7, random_state=0)``.

    No centring (docstring line 60).  Third-party dependency: scikit-learn (1.9.0 here).
    """
    from sklearn.decomposition import TruncatedSVD
    svd = TruncatedSVD(n_components=npc, n_iter=7, random_state=0)
    svd.fit(X)
    return svd.components_


def start_block(n_rows, npc, n_oversamples=10, seed=0):
    """The seeded Gaussian test matrix of ``_randomized_range_finder``
    (sklearn/utils/extmath.py: ``Q = random_state.normal(size=(A.shape[1], size))``),
    legacy MT19937 stream of ``np.random.RandomState(seed)``; ``size = npc + 10``."""
    return np.random.RandomState(seed).normal(size=(n_rows, npc + n_oversamples))


def _svd_flip_v(u, v):
    """sklearn/utils/extmath.py ``svd_flip(u, v, u_based_decision=False)``: make the
    largest-magnitude entry of every row of ``v`` positive."""
    max_abs = np.argmax(np.abs(v), axis=1)
    signs = np.sign(v[np.arange(v.shape[0]), max_abs])
    if u is not None:
        u = u * signs[np.newaxis, :]
    return u, v * signs[:, np.newaxis]


def compute_pc_restated(X, npc=1, n_iter=7, n_oversamples=10, seed=0):
    """Restatement of what sif_functions.py:65-66 executes inside scikit-learn 1.9.0:

    ``TruncatedSVD.fit_transform`` -> ``_randomized_svd(X, npc, n_iter=7,
    n_oversamples=10, power_iteration_normalizer='auto', transpose='auto',
    flip_sign=False)`` then ``svd_flip(U, VT, u_based_decision=False)``.
    'auto' normaliser = LU because n_iter > 2; 'auto' transpose = work on X.T when
    n_samples < n_features.
    """
    import scipy.linalg as sla
    M = np.asarray(X, dtype=np.float64)
    transpose = M.shape[0] < M.shape[1]
    if transpose:
        M = M.T
    k = npc + n_oversamples
    Q = start_block(M.shape[1], npc, n_oversamples, seed)
    for _ in range(n_iter):
        Q, _ = sla.lu(M @ Q, permute_l=True)
        Q, _ = sla.lu(M.T @ Q, permute_l=True)
    Q, _ = sla.qr(M @ Q, mode='economic')
    B = Q.T @ M
    Uhat, s, Vt = sla.svd(B, full_matrices=False)
    U = Q @ Uhat
    if transpose:
        U, Vt = Vt[:npc, :].T, U[:, :npc].T
    else:
        U, Vt = U[:, :npc], Vt[:npc, :]
    _, Vt = _svd_flip_v(U, Vt)
    return Vt


def compute_pc_from_gram(G, S0, npc=1, n_iter=7, transposed=False):
    """Closed form of ``compute_pc`` in the d x d Gram ``G = X^T X``.

    LU/QR normalisation preserves column spans and ``X^T (X Q) = G Q``, so sklearn's
    output is a function of ``G`` and the seeded start block only:

    * N >= d (``transposed=False``, sklearn works on ``X``): ``S0 = start_block(d, npc)``;
      with ``Q = orth(G^n_iter S0)``, ``Y = G Q``, ``T = Q^T Y`` the components are the top
      eigenvectors of ``Y T^-1 Y^T`` (the right singular vectors of ``B = Qx^T X`` with
      ``Qx = orth(X Q)``), i.e. the left singular vectors of ``W = Y T^-1/2``;
    * N < d (``transposed=True``, sklearn works on ``X^T``):
      ``S0 = X^T start_block(N, npc)`` and the components are the Rayleigh-Ritz vectors
      of ``G`` on ``span(orth(G^n_iter S0))``.

    Then rows are unit-normalised and sign-fixed by ``svd_flip(u_based_decision=False)``.
    ``csrc/pc_solve.cu`` implements exactly these steps on the device in float64.
    """
    return _pc_from_gram(G, S0, npc, n_iter, transposed=transposed)


def _orth(A):
    Q, _ = np.linalg.qr(A)
    return Q


def _pc_from_gram(G, S0, npc, n_iter, transposed):
    G = np.asarray(G, dtype=np.float64)
    Q = _orth(np.asarray(S0, dtype=np.float64))
    for _ in range(n_iter):
        Q = _orth(G @ Q)
    Y = G @ Q
    T = Q.T @ Y
    T = 0.5 * (T + T.T)
    if transposed:
        # Rayleigh-Ritz on span(Q): eigenvectors of Q^T G Q, lifted by Q.
        lam, V = np.linalg.eigh(T)
        order = np.argsort(lam)[::-1][:npc]
        comps = (Q @ V[:, order]).T
    else:
        # eigenvectors of Y T^-1 Y^T  ==  left singular vectors of W = Y T^-1/2
        lam, V = np.linalg.eigh(T)
        lam = np.maximum(lam, 0.0)
        keep = lam > lam.max() * 1e-14 if lam.max() > 0 else lam > 0
        Wm = Y @ (V[:, keep] / np.sqrt(lam[keep])[None, :])
        Sm = Wm.T @ Wm
        mu, Z = np.linalg.eigh(0.5 * (Sm + Sm.T))
        order = np.argsort(mu)[::-1][:npc]
        comps = (Wm @ Z[:, order]).T
    comps = comps / np.linalg.norm(comps, axis=1, keepdims=True)
    _, comps = _svd_flip_v(None, comps)
    return comps


def compute_pc_gram_route(X, npc=1, n_iter=7):
    """``compute_pc`` evaluated through the Gram route for a given ``X`` (float64)."""
    X = np.asarray(X, dtype=np.float64)
    n, d = X.shape
    G = X.T @ X
    if n < d:
        S0 = X.T @ start_block(n, npc)
        return _pc_from_gram(G, S0, npc, n_iter, transposed=True)
    return _pc_from_gram(G, start_block(d, npc), npc, n_iter, transposed=False)


# --------------------------------------------------------------------------- A4
def remove_pc_with(X, pc):
    """sif_functions.py:77-80 given the components: ``X - (X pc^T) * pc`` (npc == 1,
    line 78, broadcast of the (N,1) projection against the (1,d) component) or
    ``X - X pc^T pc`` (line 80)."""
    npc = pc.shape[0]
    if npc == 1:
        return X - X.dot(pc.transpose()) * pc
    return X - X.dot(pc.transpose()).dot(pc)


def remove_pc(X, npc=1):
    """sif_functions.py:69-81."""
    return remove_pc_with(X, compute_pc(X, npc))


# --------------------------------------------------------------------------- A5
def SIF_embedding(We, x, w, params, loop=False):
    """sif_functions.py:84-96: weighted average, then PC removal iff ``params.rmpc > 0``."""
    emb = (get_weighted_average_loop if loop else get_weighted_average)(We, x, w)
    if params.rmpc > 0:
        emb = remove_pc(emb, params.rmpc)
    return emb


def get_sentence_embeddings(word_embeddings, weights, text):
    """sif.py:84-94 (``RMPC = 1`` hard-coded at line 88; mask of ones at sif.py:82)."""
    text_w = seq2weight(text, np.ones(text.shape), weights)
    p = Params()
    p.rmpc = 1
    return SIF_embedding(word_embeddings, text, text_w, p)


def get_sentence_embeddings_loop(word_embeddings, weights, text):
    """sif.py:84-94 with the reference's own Python loops (the timed CPU baseline)."""
    text_w = seq2weight_loop(text, np.ones(text.shape), weights)
    p = Params()
    p.rmpc = 1
    return SIF_embedding(word_embeddings, text, text_w, p, loop=True)


def get_word_weights(word_freq_file, a=1e-3):
    """sif.py:14-32: ``a / (a + count/N)`` per word from a "word count" text file."""
    word_weights = {}
    N = 0
    with open(word_freq_file, 'r') as f:
        for line in f:
            line = line.strip()
            if len(line) > 0:
                line = line.split()
                if len(line) == 2:
                    word_weights[line[0]] = float(line[1])
                    N += float(line[1])
    for key, value in word_weights.items():
        word_weights[key] = a / (a + value / N)
    return word_weights
