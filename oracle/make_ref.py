"""Recipe for oracle/_ref/: the reference's OWN implementation of the SIF path, taken unmodified from where it lies
under /root/reference (sif_functions.py, sif.py -- pure Python, nothing to compile), so that `bench.py --impl
reference` and the `cpu_baseline` leg can time the reference's code itself instead of the oracle port.

    python oracle/make_ref.py            # no-op (exit 0) when /root/reference is not present

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but travels to the GPU box
with the snapshot, like the built .so.  Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import it; nothing under multimodal-baselines_b200/ does.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('MMB_REFERENCE_DIR', '/root/reference')
DST = os.path.join(HERE, '_ref')
FILES = ('sif_functions.py', 'sif.py')


def make():
    if not all(os.path.exists(os.path.join(REF, f)) for f in FILES):
        return None
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF, f), os.path.join(DST, f))
    with open(os.path.join(DST, 'PROVENANCE.txt'), 'w') as fh:
        fh.write('unmodified copies of %s from %s, made by oracle/make_ref.py\n' % (', '.join(FILES), REF))
    return DST


def load():
    """The reference's sif module (its get_sentence_embeddings), or None when oracle/_ref was never made."""
    if not all(os.path.exists(os.path.join(DST, f)) for f in FILES):
        return None
    import importlib.util
    mods = {}
    saved = {k: sys.modules.get(k) for k in ('sif_functions', 'sif')}
    try:
        for name in ('sif_functions', 'sif'):         # sif.py does `from sif_functions import ...`
            spec = importlib.util.spec_from_file_location(name, os.path.join(DST, name + '.py'))
            m = importlib.util.module_from_spec(spec)
            sys.modules[name] = m
            spec.loader.exec_module(m)
            mods[name] = m
    finally:
        for k, v in saved.items():                    # leave the product's modules of the same names alone
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mods['sif']


if __name__ == '__main__':
    print(make() or 'reference not present: nothing to do')
