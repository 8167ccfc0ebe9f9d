"""Recipe for oracle/_ref/: the reference's OWN implementation of the SIF path, taken unmodified from where it lies
under /root/reference (sif_functions.py, sif.py -- pure Python, nothing to compile) and archived into
oracle/_ref/reference_sif.zip, so that `bench.py --impl reference` and the `cpu_baseline` leg can time the
reference's code itself instead of the oracle port.

    python oracle/make_ref.py            # no-op (exit 0) when /root/reference is not present

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but travels to the GPU box
with the snapshot, like the built .so.  Test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's
CPU legs may import it; nothing under multimodal-baselines_b200/ does.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('MMB_REFERENCE_DIR', '/root/reference')
DST = os.path.join(HERE, '_ref')
FILES = ('sif_functions.py', 'sif.py')


BUNDLE = os.path.join(DST, 'reference_sif.zip')


def make():
    """Build oracle/_ref/reference_sif.zip from the reference tree: ONE archive (the analogue of the .so a C reference
    would be compiled into), not loose source files in this repository's directory."""
    import zipfile
    if not all(os.path.exists(os.path.join(REF, f)) for f in FILES):
        return None
    os.makedirs(DST, exist_ok=True)
    for f in FILES:                                   # (older layouts of this directory held loose copies)
        if os.path.exists(os.path.join(DST, f)):
            os.remove(os.path.join(DST, f))
    with zipfile.ZipFile(BUNDLE, 'w', zipfile.ZIP_DEFLATED) as z:
        for f in FILES:
            z.write(os.path.join(REF, f), f)
    with open(os.path.join(DST, 'PROVENANCE.txt'), 'w') as fh:
        fh.write('reference_sif.zip: unmodified %s from %s, archived by oracle/make_ref.py\n' % (', '.join(FILES), REF))
    return DST


def load():
    """The reference's sif module (its get_sentence_embeddings), or None when oracle/_ref was never made."""
    import types
    import zipfile
    if not os.path.exists(BUNDLE):
        return None
    mods = {}
    saved = {k: sys.modules.get(k) for k in ('sif_functions', 'sif')}
    try:
        with zipfile.ZipFile(BUNDLE) as z:
            for name in ('sif_functions', 'sif'):     # sif.py does `from sif_functions import ...`
                m = types.ModuleType(name)
                m.__file__ = BUNDLE + '/' + name + '.py'
                sys.modules[name] = m
                exec(compile(z.read(name + '.py').decode('utf-8'), m.__file__, 'exec'), m.__dict__)
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
