"""CUDA-graph capture with Python's cyclic collector kept out of it.

A finished loop's ``GraphedStep`` (simplesif.py) or regressor stepper (sentiment_model.py) owns
``torch.cuda.CUDAGraph`` objects and usually sits in a reference cycle (closures over ``self``), so it is
freed whenever the collector happens to run -- possibly at an allocation INSIDE the next capture.  Destroying
a graph or handing its memory pool back while a stream is capturing is an illegal call there and invalidates
the capture (``cudaErrorStreamCaptureInvalidated``, raised at some later op).  torch used to ``gc.collect()``
before every capture; since 2.6 it does so only under ``torch.compiler.config.force_cudagraph_gc``.  Keeping
the collector OFF for the duration of the capture is what matters (a full collection before every capture costs
tens of milliseconds on a heap full of torch objects -- a tenth of a grid point of the sweep); whatever is dead
is collected as usual once the capture has ended."""
import contextlib
import gc

import torch


@contextlib.contextmanager
def capture(graph, **kwargs):
    """``with capture(g): ...`` == ``with torch.cuda.graph(g): ...`` with no collection inside."""
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, **kwargs):
            yield
    finally:
        if was_enabled:
            gc.enable()
