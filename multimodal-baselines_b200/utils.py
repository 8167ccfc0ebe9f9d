"""Data helpers on either side of the hot path (reference utils.py).

``MMData`` / ``MMDataExtra`` define the batch tuple layout the MMB step consumes (reference
utils.py:193-251) and are kept as in the reference: device-resident tensors, ``__getitem__``
returns ``(idx, text, aud, vis, text_m, aud_m, vis_m, text_w[, text_a, text_a_m])``.
``normalize_data`` and ``add_positional_embeddings`` reproduce the reference's preprocessing
including its quirks (see the docstrings).  ``load_data`` and its parts read the reference's file layout
(``.npy`` ids / tables, ``word2ix`` pickle / JSON, ``.h5`` features -- the last needs ``h5py``); the datasets
themselves are external downloads of the reference.
"""
import numpy as np
import torch
from torch.utils.data import Dataset


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError as e:
        raise ImportError("the reference's feature files (data/*_data.h5, README.md:9) are HDF5: install h5py to "
                          "read them; the id / vocabulary loaders (load_word2ix, load_text_ids, "
                          "load_word_embeddings) do not need it") from e


def _read_h5_splits(path, keys):
    """``f[split][key][:]`` for the three splits (reference utils.py:35-49, 63-75, 106-120)."""
    h5py = _h5py()
    splits = ({}, {}, {})
    with h5py.File(path, 'r') as f:
        for k in keys:
            for s, name in zip(splits, ('train', 'valid', 'test')):
                s[k] = f[name][k][:]
    return splits


def load_word2ix(dataset, root='.'):
    """token -> row index of the word table: the pickle of reference utils.py:21 (MOSI) or the JSON
    mappings of utils.py:53 / 93 (POM, IEMOCAP)."""
    import json
    import os
    import pickle
    if dataset == 'mosi':
        with open(os.path.join(root, 'mosi', 'word2ix_300_mosi.pkl'), 'rb') as fh:
            return pickle.load(fh)
    if dataset in ('pom', 'iemocap'):
        with open(os.path.join(root, dataset, 'glove_mappings.%s.json' % dataset), 'r') as fh:
            return json.load(fh)
    raise ValueError('unknown dataset %r' % (dataset,))


def load_word_embeddings(dataset, root='.'):
    """The GloVe-300 table (reference utils.py:24 / 54 / 94)."""
    import os
    name = {'mosi': os.path.join('mosi', 'glove_300_mosi.npy'), 'pom': os.path.join('pom', 'glove.pom.npy'),
            'iemocap': os.path.join('iemocap', 'glove.iemocap.npy')}
    if dataset not in name:
        raise ValueError('unknown dataset %r' % (dataset,))
    return np.load(os.path.join(root, name[dataset]), allow_pickle=False)


def load_text_ids(dataset, root='.', splits=('train', 'valid', 'test')):
    """The right-padded (N, L) int64 id matrices of reference utils.py:82-88 / 122-128 (POM, IEMOCAP; MOSI
    keeps its ids inside the h5 file).  Returns one array per requested split."""
    import os
    if dataset not in ('pom', 'iemocap'):
        raise ValueError('%r has no separate id files' % (dataset,))
    return [np.load(os.path.join(root, dataset, '%s_%s_ids.npy' % (dataset, s)), allow_pickle=False) for s in splits]


def load_mosi(root='.'):
    """reference utils.py:20-50."""
    import os
    word2ix = load_word2ix('mosi', root)
    word_embeddings = load_word_embeddings('mosi', root)
    splits = _read_h5_splits(os.path.join(root, 'data', 'mosi_data.h5'),
                             ['facet', 'covarep', 'text', 'lengths', 'label', 'id'])
    return word2ix, word_embeddings, splits


def load_pom(root='.'):
    """reference utils.py:52-90."""
    import os
    word2ix = load_word2ix('pom', root)
    word_embeddings = load_word_embeddings('pom', root)
    train, valid, test = _read_h5_splits(os.path.join(root, 'data', 'pom_data.h5'), ['facet', 'covarep', 'text', 'label'])
    print(train['text'].shape)
    train['text_id'], valid['text_id'], test['text_id'] = load_text_ids('pom', root)
    print(train['text_id'].shape)
    return word2ix, word_embeddings, (train, valid, test)


def load_iemocap(args, root='.'):
    """reference utils.py:92-128."""
    import os
    word2ix = load_word2ix('iemocap', root)
    word_embeddings = load_word_embeddings('iemocap', root)
    train, valid, test = _read_h5_splits(os.path.join(root, 'data', 'iemocap_{}.h5'.format(args['emotion'])),
                                         ['facet', 'covarep', 'text', 'label'])
    print(train['text'].shape)
    train['text_id'], valid['text_id'], test['text_id'] = load_text_ids('iemocap', root)
    return word2ix, word_embeddings, (train, valid, test)


def load_data(args):
    """reference utils.py:10-18 -- ``(word2ix, word_embeddings, (train, valid, test))`` from the reference's
    file layout under the working directory (or ``args['data_root']``).  The files themselves are external
    downloads (reference README.md:9, .MISSING_LARGE_BLOBS); a missing one raises FileNotFoundError."""
    root = args.get('data_root', '.')
    if args['dataset'] == 'mosi':
        return load_mosi(root)
    elif args['dataset'] == 'pom':
        return load_pom(root)
    elif args['dataset'] == 'iemocap':
        return load_iemocap(args, root)
    else:
        raise ValueError


def add_positional_embeddings(args, data):
    """reference utils.py:130-153.

    Appends ``pos_embed_dim`` position features to ``data`` (n_points, seq_len, F).  Quirk kept
    (SURVEY.md §8d): the sin/cos transform is applied by indexing the FIRST axis
    (``idxes[2*i, :]``), i.e. only data points ``0 .. pos_embed_dim-1`` get sinusoidal features;
    every other data point gets the raw position index ``0 .. seq_len-1`` in every new column.
    """
    n_points, seq_len = data.shape[0], data.shape[1]
    pos_embed_dim = args['pos_embed_dim']
    idxes = np.tile(np.arange(seq_len, dtype=np.float32), [n_points, pos_embed_dim, 1]).transpose([0, 2, 1])
    for i in range(pos_embed_dim // 2):
        scale = 10000 ** (2 * i / pos_embed_dim)
        idxes[2 * i, :] = np.sin(idxes[2 * i, :] / scale)
        idxes[2 * i + 1, :] = np.cos(idxes[2 * i + 1, :] / scale)
    return np.concatenate([data, idxes], axis=-1)


def normalize_data(train):
    """reference utils.py:155-191: drop constant audio features, build masks from exact zeros,
    scale audio / visual features with the split's own min / max (the reference ADDS the
    minimum -- ``(x + min) * 2 / (max - min) - 1`` -- kept as is), set padding to -10."""
    audio_diff = train['covarep'].max((0, 1)) - train['covarep'].min((0, 1))
    train['covarep'] = train['covarep'][:, :, audio_diff.nonzero()[0]]
    audio_pad, vis_pad = train['covarep'] == 0, train['facet'] == 0
    audio_mask, vis_mask = (~audio_pad).astype(int), (~vis_pad).astype(int)
    audio_min, audio_max = train['covarep'].min((0, 1)), train['covarep'].max((0, 1))
    vis_min, vis_max = train['facet'].min((0, 1)), train['facet'].max((0, 1))
    train['covarep'] = (train['covarep'] + audio_min) * 2. / (audio_max - audio_min) - 1.
    train['facet'] = (train['facet'] + vis_min) * 2. / (vis_max - vis_min) - 1.
    train['covarep'][audio_pad] = -10.
    train['facet'][vis_pad] = -10.
    return train, {'covarep': audio_mask, 'facet': vis_mask}


def prep_features_device(x, pos_embed_dim=0, drop_constant=True, device=None):
    """``normalize_data`` + ``add_positional_embeddings`` + the mask extension of reference
    simplesif.py:369-375 for ONE raw feature tensor ``x`` (N, T, F), on the device (SURVEY.md §8f
    N3, libmmb_b200 ``mmb_feature_minmax`` / ``mmb_prep_features``).  Returns ``(values, mask,
    kept)``: float32 CUDA tensors of shape (N, T, F_kept + pos_embed_dim) as ``MMData`` stores them,
    and the kept source columns.  ``drop_constant`` is the reference's rule for the audio tensor
    (utils.py:163-168); the visual tensor keeps every column."""
    import _native as nv
    from _native import lib
    device = device or nv.require_cuda()
    x_t = nv.to_device(x, torch.float32, device)
    N, T, F = x_t.shape
    mn = torch.empty(F, dtype=torch.float32, device=device)
    mx = torch.empty(F, dtype=torch.float32, device=device)
    nbytes = lib.mmb_feature_minmax_workspace_bytes(N * T, F)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
    nv.check(lib.mmb_feature_minmax(nv.ptr(x_t), N * T, F, nv.ptr(mn), nv.ptr(mx), nv.ptr(ws), nbytes,
                                    nv.stream_ptr()))
    if drop_constant:
        kept = torch.nonzero(mx - mn, as_tuple=False).flatten().to(torch.int32)
    else:
        kept = torch.arange(F, dtype=torch.int32, device=device)
    F_out = int(kept.numel())
    if F_out == 0:
        raise ValueError('every feature of the tensor is constant')
    out = torch.empty((N, T, F_out + pos_embed_dim), dtype=torch.float32, device=device)
    mask = torch.empty_like(out)
    nv.check(lib.mmb_prep_features(nv.ptr(x_t), N, T, F, nv.ptr(kept.contiguous()), F_out, int(pos_embed_dim),
                                   nv.ptr(mn), nv.ptr(mx), nv.ptr(out), nv.ptr(mask), nv.stream_ptr()))
    return out, mask, kept


def normalize_data_device(train, pos_embed_dim=0, device=None):
    """Device counterpart of ``normalize_data`` (+ positional columns when ``pos_embed_dim`` > 0):
    returns ``({'covarep': tensor, 'facet': tensor}, {'covarep': mask, 'facet': mask})`` as float32
    CUDA tensors ready for ``MMData``; ``train`` is not modified."""
    cov, cov_m, _ = prep_features_device(train['covarep'], pos_embed_dim, True, device)
    fac, fac_m, _ = prep_features_device(train['facet'], pos_embed_dim, False, device)
    return {'covarep': cov, 'facet': fac}, {'covarep': cov_m, 'facet': fac_m}


def _as_f32(x, device):
    return x if torch.is_tensor(x) else torch.tensor(x, device=device, dtype=torch.float32)


class MMData(Dataset):
    """reference utils.py:193-233."""

    def __init__(self, text, audio, visual, masks, text_weights, device):
        super(Dataset, self).__init__()
        self.text = _as_f32(text, device)
        self.text_weights = _as_f32(text_weights, device)
        self.audio = _as_f32(audio, device)
        self.visual = _as_f32(visual, device)
        assert self.text.size()[0] == self.audio.size()[0]
        assert self.audio.size()[0] == self.visual.size()[0]
        assert self.text.size()[0] == self.text_weights.size()[0]
        self.text_mask = _as_f32(masks['text'], device)
        self.audio_mask = _as_f32(masks['covarep'], device)
        self.visual_mask = _as_f32(masks['facet'], device)
        self.len = self.text.size()[0]

    def __len__(self):
        return self.len

    def __getitem__(self, idx):
        return (idx, self.text[idx], self.audio[idx], self.visual[idx], self.text_mask[idx],
                self.audio_mask[idx], self.visual_mask[idx], self.text_weights[idx])


class MMDataIds(Dataset):
    """``MMData`` without the (N, L, d) word-vector tensor and its (N, L, d) mask: the text
    is kept as the (N, L) int64 ids + the table they index (``losses.TokenIds``), the text mask as the (N, L)
    ``ids != 0`` and the per-token weights as ``vocab_weights[ids]``.  ``__getitem__`` returns the same tuple
    positions as ``MMData`` (reference utils.py:231-233) with the ids in the text slot; the word term then runs
    from the ids alone."""

    def __init__(self, text_ids, audio, visual, masks, vocab_weights, table, device):
        super(Dataset, self).__init__()
        self.text_ids = torch.as_tensor(np.asarray(text_ids) if not torch.is_tensor(text_ids) else text_ids,
                                        dtype=torch.int64).to(device)
        self.table = _as_f32(table, device)
        self.audio, self.visual = _as_f32(audio, device), _as_f32(visual, device)
        self.text_weights = _as_f32(vocab_weights, device)[self.text_ids]
        self.text_mask = (self.text_ids != 0).to(torch.float32)
        self.audio_mask, self.visual_mask = _as_f32(masks['covarep'], device), _as_f32(masks['facet'], device)
        assert self.text_ids.size()[0] == self.audio.size()[0] == self.visual.size()[0]
        self.len = self.text_ids.size()[0]

    def __len__(self):
        return self.len

    def __getitem__(self, idx):
        # position 1 carries the ids (int64); the loops wrap them as losses.TokenIds(ids, self.table)
        return (idx, self.text_ids[idx], self.audio[idx], self.visual[idx],
                self.text_mask[idx], self.audio_mask[idx], self.visual_mask[idx], self.text_weights[idx])


class MMDataExtraIds(MMDataIds):
    """``MMDataExtra`` (reference utils.py:235-251) with the unaligned transcript as ids; the aligned word
    vectors that feed the Gaussian terms stay dense."""

    def __init__(self, text_ids, audio, visual, masks, vocab_weights, table, text_aligned, device):
        super(MMDataExtraIds, self).__init__(text_ids, audio, visual, masks, vocab_weights, table, device)
        self.text_aligned = _as_f32(text_aligned, device)
        self.text_aligned_mask = _as_f32(masks['text_align'], device)

    def __getitem__(self, idx):
        return MMDataIds.__getitem__(self, idx) + (self.text_aligned[idx], self.text_aligned_mask[idx])


class MMDataExtra(MMData):
    """reference utils.py:235-251 -- adds the aligned text vectors + mask (POM / IEMOCAP)."""

    def __init__(self, text, audio, visual, masks, text_weights, text_aligned, device):
        super(MMDataExtra, self).__init__(text, audio, visual, masks, text_weights, device)
        self.text_aligned = _as_f32(text_aligned, device)
        self.text_aligned_mask = _as_f32(masks['text_align'], device)

    def __getitem__(self, idx):
        return MMData.__getitem__(self, idx) + (self.text_aligned[idx], self.text_aligned_mask[idx])
