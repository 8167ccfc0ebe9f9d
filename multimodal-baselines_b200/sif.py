"""Drop-in for the reference's ``sif.py``: word weights and
``get_sentence_embeddings``, the call ``simplesif.main`` makes per split (reference
simplesif.py:296-311).  The embedding itself runs through ``mmb_sif_embedding_host`` /
``mmb_sif_embedding`` of libmmb_b200.so.
"""
import ctypes as C
import os

import numpy as np
import torch

import _native as nv
from _native import lib
from sif_functions import Params, seq2weight, SIF_embedding, start_block, sif_embedding_device  # noqa: F401

"""
1. Initialize sentence embedding using the SIF algorithm over training data

word_weights : a / (a + p(w)) - calculate unigram probabilities over training data
"""


def get_word_weights(word_freq_file, a=1e-3):
    """reference sif.py:14-32 -- ``a / (a + count/N)`` from a "word count" text file."""
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
                else:
                    print(line)
    for key, value in word_weights.items():
        word_weights[key] = a / (a + value / N)
    return word_weights


def load_weights(args):
    """reference sif.py:34-42."""
    if args['dataset'] == 'mosi':
        return load_mosi_weights()
    elif args['dataset'] == 'pom':
        return load_pom_weights()
    elif args['dataset'] == 'iemocap':
        return load_iemocap_weights()
    else:
        raise NotImplementedError


def load_pom_weights():
    """reference sif.py:44-47."""
    weights = np.load('pom/pom_word_weights.npy').squeeze()
    print(weights.shape)
    return weights


def load_iemocap_weights():
    """reference sif.py:49-52."""
    weights = np.load('iemocap/iemocap_word_weights.npy').squeeze()
    print(weights.shape)
    return weights


def load_mosi_weights(word2ix=None):
    """reference sif.py:54-76.  The reference's regeneration branch reads an undefined
    global ``word2ix`` (line 63, NameError); here it is an optional argument."""
    if os.path.isfile('word_weights.npy'):
        return np.load('word_weights.npy', allow_pickle=False).squeeze()
    if word2ix is None:
        raise NameError("name 'word2ix' is not defined")   # what the reference does here
    word_weights = get_word_weights('SIF/auxiliary_data/enwiki_vocab_min200.txt')
    weights = np.zeros((max(word2ix.values()) + 1))
    unk = 0
    for word, ix in word2ix.items():
        if word.lower() not in word_weights.keys():
            weights[ix] = 1.
            unk += 1
        else:
            weights[ix] = word_weights[word.lower()]
    print("# of words with unknown weight", unk)
    print(weights[:5])
    np.save('word_weights.npy', weights, allow_pickle=False)
    return weights


def get_sentence_word_weights(text, weights):
    """reference sif.py:78-82 -- weights for each word in each sentence (mask of ones)."""
    return seq2weight(text, np.ones(np.shape(text)), weights)


def get_sentence_embeddings(word_embeddings, weights, text, out=None):
    """reference sif.py:84-94 -- SIF embedding of every utterance with the first principal
    component removed (``RMPC = 1``, line 88).

    NumPy in -> float64 NumPy out, as the reference; CUDA tensors in -> float32 CUDA tensor
    out (device-resident fused path).  ``out`` may be a preallocated (N, d) float64/float32
    host array (e.g. ``_native.PinnedArray(...).array``) to receive the result.
    """
    RMPC = 1
    dev = nv.require_cuda()
    if isinstance(text, torch.Tensor) and text.is_cuda:
        table_t = nv.to_device(word_embeddings, torch.float32, dev)
        w_t = nv.to_device(weights, torch.float32, dev).reshape(-1)
        return sif_embedding_device(table_t, w_t, text.contiguous(), npc=RMPC)

    ids = np.ascontiguousarray(np.asarray(text), dtype=np.int64)
    if ids.ndim != 2:
        raise ValueError('text must be (n_samples, seq_len)')
    n, L = ids.shape
    table_t = nv.to_device(word_embeddings, torch.float32, dev)
    w_t = nv.to_device(np.asarray(weights).reshape(-1), torch.float32, dev)
    V, d = table_t.shape
    if w_t.numel() < V:
        if ids.size and ids.max() >= w_t.numel():
            raise IndexError('index %d is out of bounds for axis 0 with size %d' % (ids.max(), w_t.numel()))
        w_t = torch.cat([w_t, w_t.new_zeros(V - w_t.numel())])
    if out is None:
        out = np.empty((n, d), dtype=np.float64)
    assert out.shape == (n, d) and out.flags['C_CONTIGUOUS'] and out.dtype in (np.float64, np.float32)
    omega = np.ascontiguousarray(start_block(d if n >= d else n, RMPC))
    torch.cuda.current_stream().synchronize()      # table / weights uploads are done
    nv.check(lib.mmb_sif_embedding_host(nv.ptr(table_t), V, d, nv.ptr(w_t), nv.np_ptr(ids), n, L, RMPC,
                                        nv.np_ptr(omega), nv.np_ptr(out), int(out.dtype == np.float64),
                                        None, nv.GRAM_AUTO, 0))
    return out
