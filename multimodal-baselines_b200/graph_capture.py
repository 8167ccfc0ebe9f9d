"""CUDA-graph capture with Python's cyclic collector kept out of it.

A finished loop's ``GraphedStep`` (simplesif.py) or regressor stepper (sentiment_model.py) owns
``torch.cuda.CUDAGraph`` objects and usually sits in a reference cycle (closures over ``self``), so it is
freed whenever the collector happens to run -- possibly at an allocation INSIDE the next capture.  Destroying
a graph or handing its memory pool back while a stream is capturing is an illegal call there and invalidates
the capture (``cudaErrorStreamCaptureInvalidated``, raised at some later op).  torch used to ``gc.collect()``
before every capture; since 2.6 it does so only under ``torch.compiler.config.force_cudagraph_gc``.  So:
collect before the capture, and keep the collector off until it ends."""
import contextlib
import gc

import torch


@contextlib.contextmanager
def capture(graph, **kwargs):
    """``with capture(g): ...`` == ``with torch.cuda.graph(g): ...`` with no collection inside."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, **kwargs):
            yield
    finally:
        if was_enabled:
            gc.enable()
