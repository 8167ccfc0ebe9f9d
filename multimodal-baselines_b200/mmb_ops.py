"""torch.autograd bindings of the MMB step kernels.

Each Function is one call into libmmb_b200.so that produces the value together with the
analytic gradients (Appendix A.3-A.5), so ``backward`` is a broadcast multiply (Gaussian,
word term) or one more library call (heads).  Inputs are made float32 / contiguous CUDA
tensors; anything else is an error -- there is no CPU fallback.
"""
import ctypes as C
import weakref

import torch

import _native as nv
from _native import lib


def _f32(t):
    if not t.is_cuda:
        raise nv.MMBError('libmmb_b200 needs CUDA tensors (got %s); there is no CPU fallback' % t.device)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[0 if t is None else t.data_ptr() for t in tensors])


def _int_array(vals):
    return (C.c_int * len(vals))(*[int(v) for v in vals])


def scale_multi(ins, rows=None, elems=None):
    """out[i] = ins[i] * rows[i][:, None] * elems[i] for up to 16 (B, D_i) tensors in ONE launch
    (``mmb_scale_multi``); ``rows`` / ``elems`` entries may be None."""
    B = ins[0].shape[0]
    outs = [torch.empty_like(t) for t in ins]
    for lo in range(0, len(ins), 16):
        sl = slice(lo, lo + 16)
        nv.check(lib.mmb_scale_multi(len(ins[sl]), B, _int_array([t.shape[1] for t in ins[sl]]), _ptr_array(ins[sl]),
                                     _ptr_array(rows[sl]) if rows is not None else None,
                                     _ptr_array(elems[sl]) if elems is not None else None,
                                     _ptr_array(outs[sl]), nv.stream_ptr()))
    return outs


def gather_multi(srcs, idx):
    """[s[idx] for s in srcs] for float32 CUDA tensors sharing one int64 index vector, in ONE launch
    (``mmb_gather_multi``): the batch tuple of ``MMData.__getitem__``."""
    B = int(idx.shape[0])
    srcs = [_f32(s) for s in srcs]
    outs = [torch.empty((B,) + tuple(s.shape[1:]), dtype=torch.float32, device=s.device) for s in srcs]
    W = (C.c_int64 * len(srcs))(*[int(s[0].numel()) for s in srcs])
    nv.check(lib.mmb_gather_multi(len(srcs), B, nv.ptr(idx.contiguous()), _ptr_array(srcs), W, _ptr_array(outs),
                                  nv.stream_ptr()))
    return outs


class HeadsFunction(torch.autograd.Function):
    """All (mu, log_sigma) heads of AudioVisualGeneratorMultimodal in one launch
    (reference models.py:196-202).  ``apply(z, is_log_sigma, W0, b0, W1, b1, ...)`` returns
    one (B, D_h) tensor per head; heads flagged in ``is_log_sigma`` get the exp() epilogue."""

    @staticmethod
    def forward(ctx, z, is_log_sigma, *params):
        z = _f32(z)
        Ws = [_f32(p) for p in params[0::2]]
        bs = [_f32(p) for p in params[1::2]]
        B, d = z.shape
        Ds = [w.shape[0] for w in Ws]
        outs = [torch.empty((B, D), dtype=torch.float32, device=z.device) for D in Ds]
        nv.check(lib.mmb_heads_forward(nv.ptr(z), B, d, len(Ws), _ptr_array(Ws), _ptr_array(bs),
                                       _int_array(Ds), _int_array(is_log_sigma), _ptr_array(outs),
                                       nv.stream_ptr()))
        ctx.is_log_sigma = list(is_log_sigma)
        ctx.Ds = Ds
        ctx.save_for_backward(z, *Ws, *[o for o, ls in zip(outs, is_log_sigma) if ls])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        saved = ctx.saved_tensors
        n = len(ctx.Ds)
        z, Ws, sig = saved[0], saved[1:1 + n], list(saved[1 + n:])
        B, d = z.shape
        gpre = [torch.zeros((B, ctx.Ds[h]), dtype=torch.float32, device=z.device) if gouts[h] is None
                else _f32(gouts[h]) for h in range(n)]
        ls = [h for h in range(n) if ctx.is_log_sigma[h]]
        if ls:                                # d sigma / d s = sigma  (sigma = exp(s)): one launch
            scaled = scale_multi([gpre[h] for h in ls], None, list(sig))
            for h, g in zip(ls, scaled):
                gpre[h] = g
        need_z = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[2:])
        dz = torch.empty_like(z) if need_z else None
        dWs = dbs = flat = None
        if need_w:
            # every head's dW and db are views of ONE flat buffer, in parameter order (W0, b0, W1, b1, ...): the
            # data-parallel step sums it over the ranks with one exchange right behind the kernel that fills it
            sizes = []
            for w, D in zip(Ws, ctx.Ds):
                sizes.extend([w.numel(), D])
            flat = torch.empty(sum(sizes), dtype=torch.float32, device=z.device)
            views, off = [], 0
            for n_el in sizes:
                views.append(flat[off:off + n_el])
                off += n_el
            dWs = [v.view_as(w) for v, w in zip(views[0::2], Ws)]
            dbs = views[1::2]
        Ds_c = _int_array(ctx.Ds)
        nbytes = lib.mmb_heads_backward_workspace_bytes(B, d, n, Ds_c) if need_z else 0
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=z.device)
        nv.check(lib.mmb_heads_backward(nv.ptr(z), B, d, n, _ptr_array(Ws), Ds_c, _ptr_array(gpre),
                                        nv.ptr(dz), _ptr_array(dWs) if need_w else None,
                                        _ptr_array(dbs) if need_w else None, nv.ptr(ws), nbytes,
                                        nv.stream_ptr()))
        if need_w:
            import mmb_dp
            dp = mmb_dp.active()
            if dp is not None and getattr(dp, 'reduce_head_grads', False):
                dp.allreduce_(flat)          # sum over ranks of the local-batch sums (already scaled by 1 / B_global)
        grads = [dz, None]
        for h in range(n):
            grads.append(dWs[h] if need_w and ctx.needs_input_grad[2 + 2 * h] else None)
            grads.append(dbs[h] if need_w and ctx.needs_input_grad[3 + 2 * h] else None)
        return tuple(grads)


class GaussLLFunction(torch.autograd.Function):
    """Masked diagonal-Gaussian log-likelihood of every modality in one launch (reference
    losses.py:13-34, 251-256).  ``apply(segments, status, mu_0, sigma_0, mu_1, sigma_1, ...)``
    with ``segments[m]`` = list of (values, mask) pairs whose feature axes concatenate to
    modality m (the reference's torch.cat, simplesif.py:94-113, without materialising it);
    returns lp of shape (n_mod, B)."""

    @staticmethod
    def forward(ctx, segments, status, *mu_sigma):
        mus = [_f32(t) for t in mu_sigma[0::2]]
        sigmas = [_f32(t) for t in mu_sigma[1::2]]
        n_mod = len(mus)
        B = mus[0].shape[0]
        vals, masks, Fs, n_seg = [], [], [], []
        T = None
        for m in range(n_mod):
            n_seg.append(len(segments[m]))
            D = 0
            for v, k in segments[m]:
                v, k = _f32(v), _f32(k)
                if v.dim() != 3 or v.shape != k.shape or v.shape[0] != B:
                    raise ValueError('values/mask must both be (batch, seq_len, n_features)')
                T = v.shape[1] if T is None else T
                if v.shape[1] != T:
                    raise ValueError('all modalities must share seq_len in one call')
                vals.append(v)
                masks.append(k)
                Fs.append(v.shape[2])
                D += v.shape[2]
            if mus[m].shape != (B, D) or sigmas[m].shape != (B, D):
                raise RuntimeError('The size of mu/sigma %s must match the data (%d, %d)'
                                   % (tuple(mus[m].shape), B, D))
        lp = torch.empty((n_mod, B), dtype=torch.float32, device=mus[0].device)
        dmu = [torch.empty_like(t) for t in mus]
        dsg = [torch.empty_like(t) for t in sigmas]
        nv.check(lib.mmb_gauss_ll(B, T, n_mod, _int_array(n_seg), _ptr_array(vals), _ptr_array(masks),
                                  _int_array(Fs), _ptr_array(mus), _ptr_array(sigmas), nv.ptr(lp),
                                  _ptr_array(dmu), _ptr_array(dsg), nv.ptr(status), nv.stream_ptr()))
        ctx.save_for_backward(*dmu, *dsg)
        ctx.n_mod = n_mod
        return lp

    @staticmethod
    def backward(ctx, g):
        n = ctx.n_mod
        dmu, dsg = ctx.saved_tensors[:n], ctx.saved_tensors[n:]
        g = _f32(g)
        scaled = scale_multi(list(dmu) + list(dsg), [g[m] for m in range(n)] * 2, None)   # chain rule, one launch
        grads = [None, None]
        for m in range(n):
            grads.append(scaled[m] if ctx.needs_input_grad[2 + 2 * m] else None)
            grads.append(scaled[n + m] if ctx.needs_input_grad[3 + 2 * m] else None)
        return tuple(grads)


def gauss_moments(values, mask):
    """(N, T, F) values + float mask -> (N, 3, F) masked moments over time [S0 | mean | M2]
    (``mmb_gauss_moments``): computed once per dataset, see GaussLLStatsFunction."""
    v, k = _f32(values), _f32(mask)
    if v.dim() != 3 or v.shape != k.shape:
        raise ValueError('values/mask must both be (n, seq_len, n_features)')
    N, T, F = v.shape
    stats = torch.empty((N, 3, F), dtype=torch.float32, device=v.device)
    nv.check(lib.mmb_gauss_moments(nv.ptr(v), nv.ptr(k), N, T, F, nv.ptr(stats), nv.stream_ptr()))
    return stats


class GaussLLStatsFunction(torch.autograd.Function):
    """GaussLLFunction fed with the per-utterance moments of the batch rows instead of the (B, T, F) values
    and masks (``mmb_gauss_ll_stats``).  ``segments[m]`` = list of (B, 3, F) tensors."""

    @staticmethod
    def forward(ctx, segments, status, *mu_sigma):
        mus = [_f32(t) for t in mu_sigma[0::2]]
        sigmas = [_f32(t) for t in mu_sigma[1::2]]
        n_mod = len(mus)
        B = mus[0].shape[0]
        stats, Fs, n_seg = [], [], []
        for m in range(n_mod):
            n_seg.append(len(segments[m]))
            D = 0
            for st in segments[m]:
                st = _f32(st)
                if st.dim() != 3 or st.shape[0] != B or st.shape[1] != 3:
                    raise ValueError('moments must be (batch, 3, n_features)')
                stats.append(st)
                Fs.append(st.shape[2])
                D += st.shape[2]
            if mus[m].shape != (B, D) or sigmas[m].shape != (B, D):
                raise RuntimeError('The size of mu/sigma %s must match the data (%d, %d)'
                                   % (tuple(mus[m].shape), B, D))
        lp = torch.empty((n_mod, B), dtype=torch.float32, device=mus[0].device)
        dmu = [torch.empty_like(t) for t in mus]
        dsg = [torch.empty_like(t) for t in sigmas]
        nv.check(lib.mmb_gauss_ll_stats(B, n_mod, _int_array(n_seg), _ptr_array(stats), _int_array(Fs),
                                        _ptr_array(mus), _ptr_array(sigmas), nv.ptr(lp), _ptr_array(dmu),
                                        _ptr_array(dsg), nv.ptr(status), nv.stream_ptr()))
        ctx.save_for_backward(*dmu, *dsg)
        ctx.n_mod = n_mod
        return lp

    backward = staticmethod(GaussLLFunction.backward)


class CombineLPFunction(torch.autograd.Function):
    """``other_w * lp.sum(0) + word_w * wlp`` (reference losses.py:267-272) with its backward, one launch each
    way (``mmb_combine_lp``).  lp (M, B), wlp (B,) -> (B,)."""

    @staticmethod
    def forward(ctx, lp, wlp, other_w, word_w):
        """``other_w`` / ``word_w``: Python numbers, or one-element float32 CUDA tensors -- the kernel then reads the
        weights from device memory, so a captured graph serves every grid point's weights (sweep.py)."""
        lp, wlp = _f32(lp), _f32(wlp)
        M, B = lp.shape
        out = torch.empty(B, dtype=torch.float32, device=lp.device)
        dev_w = torch.is_tensor(other_w)
        ow, ww = (0., 0.) if dev_w else (float(other_w), float(word_w))
        nv.check(lib.mmb_combine_lp(nv.ptr(lp), nv.ptr(wlp), M, B, ow, ww, nv.ptr(other_w) if dev_w else None,
                                    nv.ptr(word_w) if dev_w else None, nv.ptr(out), nv.stream_ptr()))
        ctx.w = (ow, ww, M, B, other_w if dev_w else None, word_w if dev_w else None)
        return out

    @staticmethod
    def backward(ctx, g):
        ow, ww, M, B, ow_t, ww_t = ctx.w
        g = _f32(g)
        g_lp = torch.empty((M, B), dtype=torch.float32, device=g.device)
        g_wlp = torch.empty(B, dtype=torch.float32, device=g.device)
        nv.check(lib.mmb_combine_lp_backward(nv.ptr(g), M, B, ow, ww, nv.ptr(ow_t), nv.ptr(ww_t), nv.ptr(g_lp),
                                             nv.ptr(g_wlp), nv.stream_ptr()))
        return g_lp, g_wlp, None, None


class WordLLFunction(torch.autograd.Function):
    """Angular word log-probability (reference losses.py:68-95) with its gradient w.r.t. the
    latents; the word table, token vectors, weights and mask are constants of the step."""

    @staticmethod
    def forward(ctx, latents, table, word_w, sent, mask, a, status):
        e = _f32(latents)
        table = _f32(table)
        word_w = _f32(word_w)
        sent = sent if (sent.dtype == torch.float32 and sent.is_cuda and sent.stride(-1) == 1) else _f32(sent)
        if not (mask.dtype == torch.float32 and mask.is_cuda):
            mask = _f32(mask)
        B, d = e.shape
        V = table.shape[0]
        L = word_w.shape[1]
        if sent.shape[:2] != (B, L) or sent.shape[2] != d or mask.shape[:2] != (B, L):
            raise RuntimeError('word term: shapes do not match (latents %s, sent %s, weights %s, mask %s)'
                               % (tuple(e.shape), tuple(sent.shape), tuple(word_w.shape), tuple(mask.shape)))
        inv_norm = table_inv_norm(table)
        lp = torch.empty(B, dtype=torch.float32, device=e.device)
        grad = torch.empty_like(e)
        nbytes = lib.mmb_word_ll_workspace_bytes(B, V, d)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=e.device)
        nv.check(lib.mmb_word_ll(nv.ptr(e), B, d, nv.ptr(table), nv.ptr(inv_norm), V,
                                 C.c_void_p(sent.data_ptr()), sent.stride(0), sent.stride(1),
                                 nv.ptr(word_w), C.c_void_p(mask.data_ptr()), mask.stride(0), mask.stride(1),
                                 L, float(a), nv.ptr(lp), nv.ptr(grad), nv.ptr(ws), nbytes, nv.ptr(status),
                                 nv.stream_ptr()))
        ctx.save_for_backward(grad)
        return lp

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return g.unsqueeze(1) * grad, None, None, None, None, None, None


class WordLLIdsFunction(torch.autograd.Function):
    """The same term when the token vectors are rows of the word table: ids (B, L) int64 instead of the
    (B, L, d) vectors (``mmb_word_ll_ids``).  ``mask`` is (B, L) float or None (= ids != 0)."""

    @staticmethod
    def forward(ctx, latents, table, word_w, ids, mask, a, status):
        e = _f32(latents)
        table = _f32(table)
        word_w = _f32(word_w)
        if not ids.is_cuda or ids.dtype != torch.int64:
            raise nv.MMBError('token ids must be an int64 CUDA tensor')
        if ids.stride(-1) != 1:
            ids = ids.contiguous()
        B, d = e.shape
        V = table.shape[0]
        L = word_w.shape[1]
        if tuple(ids.shape) != (B, L) or (mask is not None and tuple(mask.shape[:2]) != (B, L)):
            raise RuntimeError('word term: shapes do not match (latents %s, ids %s, weights %s)'
                               % (tuple(e.shape), tuple(ids.shape), tuple(word_w.shape)))
        if mask is not None and not (mask.dtype == torch.float32 and mask.is_cuda):
            mask = _f32(mask)
        inv_norm = table_inv_norm(table)
        lp = torch.empty(B, dtype=torch.float32, device=e.device)
        grad = torch.empty_like(e)
        nbytes = lib.mmb_word_ll_ids_workspace_bytes(B, V, d)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=e.device)
        nv.check(lib.mmb_word_ll_ids(nv.ptr(e), B, d, nv.ptr(table), nv.ptr(inv_norm), V, nv.ptr(ids), ids.stride(0),
                                     nv.ptr(word_w), C.c_void_p(mask.data_ptr()) if mask is not None else None,
                                     mask.stride(0) if mask is not None else 0,
                                     mask.stride(1) if mask is not None else 0, L, float(a), nv.ptr(lp),
                                     nv.ptr(grad), nv.ptr(ws), nbytes, nv.ptr(status), nv.stream_ptr()))
        ctx.save_for_backward(grad)
        return lp

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return g.unsqueeze(1) * grad, None, None, None, None, None, None


_STATUS_SINK = None
_INV_NORM_CACHE = {}


def set_status_sink(t):
    """Route every kernel's status bits into ONE persistent device word `t` (int32, shape (1,))
    instead of a fresh word per call, and stop the per-call host read-back: the caller checks
    the sink when it next synchronises (once per epoch in the graph-captured loop).  ``None``
    restores the per-call behaviour.  Returns the previous sink."""
    global _STATUS_SINK
    prev, _STATUS_SINK = _STATUS_SINK, t
    return prev


def status_deferred():
    return _STATUS_SINK is not None


def new_status(device):
    if _STATUS_SINK is not None:
        return _STATUS_SINK
    return torch.zeros(1, dtype=torch.int32, device=device)


def table_inv_norm(table):
    """1 / max(||row||, 1e-8) of the word table (torch CosineSimilarity's clamp), cached per table
    TENSOR OBJECT (weak reference + ``_version``) -- the table is a constant of the optimisation loop.
    The key is the object, not its address: a new table of the same shape that the caching allocator
    places at a freed table's address must not inherit the old norms.  Several tables can be cached
    at once; an entry dies with its table.  Callers that bake the returned buffer into a CUDA graph
    must keep their own reference to it (``GraphedStep`` does)."""
    key = id(table)
    hit = _INV_NORM_CACHE.get(key)
    if hit is not None and hit[0]() is table and hit[1] == table._version:
        return hit[2]
    V, d = table.shape
    inv_norm = torch.empty(V, dtype=torch.float32, device=table.device)
    nv.check(lib.mmb_row_inv_norm(nv.ptr(table), V, d, nv.ptr(inv_norm), nv.stream_ptr()))
    if not torch.cuda.is_current_stream_capturing():
        ref = weakref.ref(table, lambda _r, k=key: _INV_NORM_CACHE.pop(k, None))
        _INV_NORM_CACHE[key] = (ref, table._version, inv_norm)
    return inv_norm


def cached_inv_norms():
    """The inverse-norm buffers currently cached (GraphedStep keeps them alive for as long as its graphs,
    which have their addresses baked in, can be replayed)."""
    return [hit[2] for hit in _INV_NORM_CACHE.values()]
