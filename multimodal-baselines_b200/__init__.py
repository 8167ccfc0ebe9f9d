"""B200-native SIF / MMB utterance-embedding hot path of yaochie/multimodal-baselines.

The reference's modules are top-level (``import sif_functions``, ``import losses`` ...), so
this directory is meant to be put on ``sys.path`` in place of the reference checkout; doing
``importlib.import_module('multimodal-baselines_b200')`` does exactly that.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)
