"""Golden generation only: import the reference's utils.py (MMData) with h5py stubbed
(h5py is not installed here and MMData does not use it)."""
import sys
import types

if 'h5py' not in sys.modules:
    try:
        import h5py  # noqa: F401
    except ImportError:
        sys.modules['h5py'] = types.ModuleType('h5py')
import importlib.util
import os

_ref = os.environ.get('MMB_REFERENCE', '/root/reference')
_spec = importlib.util.spec_from_file_location('ref_utils', os.path.join(_ref, 'utils.py'))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
MMData = _mod.MMData
MMDataExtra = _mod.MMDataExtra
add_positional_embeddings = _mod.add_positional_embeddings
normalize_data = _mod.normalize_data
