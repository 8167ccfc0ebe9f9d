"""Latency of the SIF path at the reference's real sizes (MOSI- and POM-sized
inputs are launch-latency bound, so they are reported as latency, not roofline fractions):
``sif.get_sentence_embeddings`` (NumPy in, float64 NumPy out, the drop-in call) and the
device-resident pipeline, against the CPU oracle port of the reference on the host cores.

    python tools/bench_small.py            # one JSON line per shape
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import numpy as np
import torch
import sif
import sif_functions as sf
from oracle import sif_oracle as so

SHAPES = {'mosi_all (2199 x 20, V 3016)': (2199, 20, 3016), 'pom_test (203 x 1357, V 7763)': (203, 1357, 7763),
          'pom_train_like (600 x 1357, V 7763)': (600, 1357, 7763)}


def inputs(n, L, V, seed=0):
    rng = np.random.default_rng(seed)
    We = (0.4 * rng.standard_normal((V, 300)) + 0.3 * rng.standard_normal((1, 300))).astype(np.float32)
    We[0] = 0
    p = 1.0 / np.arange(1, V) ** 1.1
    p /= p.sum()
    ids = rng.choice(np.arange(1, V), size=(n, L), p=p).astype(np.int64)
    lens = rng.integers(max(1, L // 8), L + 1, size=n)
    ids[np.arange(L)[None, :] >= lens[:, None]] = 0
    weights = np.concatenate([[1.0], 1e-3 / (1e-3 + p)])
    return We, weights, ids


def best_of(fn, reps):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), float(np.median(ts))


for name, (n, L, V) in SHAPES.items():
    We, weights, ids = inputs(n, L, V)
    got = sif.get_sentence_embeddings(We, weights, ids)          # warm-up + parity
    want = so.get_sentence_embeddings(We, weights, ids)
    err = float(np.abs(got - want).max() / np.abs(want).max())
    api_best, api_med = best_of(lambda: sif.get_sentence_embeddings(We, weights, ids), 10)
    dev = torch.device('cuda')
    t_We, t_w, t_ids = torch.tensor(We, device=dev), torch.tensor(weights, device=dev, dtype=torch.float32), \
        torch.tensor(ids, device=dev)
    sf.sif_embedding_device(t_We, t_w, t_ids, npc=1)
    dev_best, dev_med = best_of(lambda: sf.sif_embedding_device(t_We, t_w, t_ids, npc=1, check=False), 20)
    t0 = time.perf_counter()
    so.get_sentence_embeddings_loop(We, weights, ids)
    cpu = time.perf_counter() - t0
    print(json.dumps({'shape': name, 'max_rel_err_vs_oracle': err,
                      'api_numpy_in_out_ms': {'best': api_best * 1e3, 'median': api_med * 1e3},
                      'device_resident_ms': {'best': dev_best * 1e3, 'median': dev_med * 1e3},
                      'cpu_port_ms': cpu * 1e3, 'cpu_cores': len(os.sched_getaffinity(0)),
                      'utt_per_s_device': n / dev_best, 'utt_per_s_cpu_port': n / cpu}), flush=True)
