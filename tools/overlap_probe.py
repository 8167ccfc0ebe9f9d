"""GPU probe of mmb_sif_embed_gram: the Gram of chunk c beside the embed of chunk c + 1 (csrc/api.cu) against the two
stages run one after the other, on the bench workload.  Prints one JSON line per setting:
   python tools/overlap_probe.py [N] [sms,sms,...] [chunks,chunks,...] [grid,grid,...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import numpy as np
import torch
import bench
import _native as nv
from _native import lib


def run(table, vw, ids, emb, G, ws, ws_bytes, sms, chunks, grid, reps=5, warm=2):
    nv.check(lib.mmb_set_option(b'overlap_sms', int(sms)))
    nv.check(lib.mmb_set_option(b'overlap_chunks', int(chunks)))
    nv.check(lib.mmb_set_option(b'overlap_grid', int(grid)))
    n, L = ids.shape
    V, d = table.shape
    st = torch.zeros(1, dtype=torch.int32, device=ids.device)
    ms = []
    for it in range(warm + reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nv.check(lib.mmb_sif_embed_gram(nv.ptr(table), V, d, nv.ptr(vw), nv.ptr(ids), n, L, nv.ptr(emb), nv.ptr(st),
                                        nv.ptr(G), nv.ptr(ws), ws_bytes, 0, nv.stream_ptr()))
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    assert int(st.item()) == 0
    return float(np.mean(ms[warm:])), float(np.min(ms[warm:]))


def ints(s):
    return [int(v) for v in s.split(',')]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    sms_list = ints(sys.argv[2]) if len(sys.argv) > 2 else [24, 32, 40, 48]
    chunk_list = ints(sys.argv[3]) if len(sys.argv) > 3 else [10]
    grid_list = ints(sys.argv[4]) if len(sys.argv) > 4 else [32]
    dev = torch.device('cuda')
    table, vw, p = bench.make_table_and_weights(dev)
    ids = bench.make_ids(dev, n, bench.L_TOK, p, seed=1000)
    V, d = table.shape
    emb = torch.empty((n, d), dtype=torch.float32, device=dev)
    G = torch.empty((d, d), dtype=torch.float32, device=dev)
    ws_bytes = lib.mmb_sif_embed_gram_workspace_bytes(V, d, n, bench.L_TOK, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    ms0, best0 = run(table, vw, ids, emb, G, ws, ws_bytes, 0, 10, 8)
    emb0, G0 = emb[:200_000].clone(), G.double().clone()
    tail0 = emb[-200_000:].clone()
    print(json.dumps({'setting': 'serial', 'n': n, 'ms': ms0, 'best_ms': best0}), flush=True)
    for chunks in chunk_list:
        for grid in grid_list:
            for sms in sms_list:
                emb.zero_()
                G.zero_()
                ms, best = run(table, vw, ids, emb, G, ws, ws_bytes, sms, chunks, grid)
                same = bool(torch.equal(emb[:200_000], emb0) and torch.equal(emb[-200_000:], tail0))
                gerr = float((G.double() - G0).abs().max() / G0.abs().max())
                print(json.dumps({'setting': 'overlap', 'gram_sms': sms, 'chunks': chunks, 'embed_ctas_per_sm': grid,
                                  'ms': ms, 'best_ms': best, 'speedup': ms0 / ms, 'emb_bits_equal': same,
                                  'gram_rel_diff': gerr}), flush=True)


if __name__ == '__main__':
    main()
