"""CPU: libmmb_b200.so builds, loads, and exports every symbol include/mmb_b200.h declares;
the Python binding table mirrors the header; the product path refuses to run without CUDA."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'mmb_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'MMB_API\s+[\w\s\*]+?\b(mmb_\w+)\s*\(', text)))


@pytest.fixture(scope='module')
def built_lib():
    import __graft_entry__ as g
    g.build()
    import _native
    return _native


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for need in ('mmb_seq2weight', 'mmb_weighted_average', 'mmb_sif_embed', 'mmb_gram',
                 'mmb_pc_from_gram', 'mmb_remove_pc', 'mmb_sif_embedding', 'mmb_sif_embedding_host',
                 'mmb_heads_forward', 'mmb_gauss_ll', 'mmb_word_ll', 'mmb_last_error'):
        assert need in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.mmb_version() >= 100


def test_binding_table_matches_header(built_lib):
    assert sorted(built_lib.SIGNATURES) == declared_symbols()
    text = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    for name, (_res, args) in built_lib.SIGNATURES.items():
        m = re.search(r'\b%s\s*\(([^;]*?)\)\s*;' % name, text, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ('', 'void') else params.count(',') + 1
        assert n == len(args), (name, n, len(args))


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import numpy as np
    import sif_functions
    import sif
    with pytest.raises(built_lib.MMBError):
        sif_functions.get_weighted_average(np.zeros((4, 8), np.float32), np.zeros((2, 3), np.int64),
                                           np.ones((2, 3), np.float32))
    with pytest.raises(built_lib.MMBError):
        sif.get_sentence_embeddings(np.zeros((4, 8), np.float32), np.ones(4), np.zeros((2, 3), np.int64))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'multimodal-baselines_b200')
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f
                assert '/root/reference' not in src, f
