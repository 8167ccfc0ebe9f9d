"""GPU probe of the tcgen05 Gram kernel's MMA issue order (csrc/gram_tc.cu, MMB_TC_COLLECT): time and error against the
float64 Gram for each mode.   python tools/gram_probe.py [N] [modes]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import numpy as np
import torch
import _native as nv
from _native import lib


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    # modes: collect[:seg[:passes[:opt]]] , ...   (the committed r02_gram_probe.jsonl also has 'dbg' rows from a build
    # with ablation switches: 1 = no block C, 2 = no split, 4 = no MMAs -- timing only, since removed)
    modes = [tuple(int(x) for x in v.split(':')) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [(0, 64, 3, 0), (1, 64, 3, 15), (0, 64, 3, 0), (1, 64, 3, 15)]
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(5)
    d = 300
    X = torch.randn((n, d), generator=g, device=dev) * 0.05 + torch.randn((1, d), generator=g, device=dev) * 0.02
    want = torch.zeros((d, d), dtype=torch.float64, device=dev)
    for r0 in range(0, n, 1 << 18):
        xb = X[r0:r0 + (1 << 18)].double()
        want += xb.T @ xb
    G = torch.empty((d, d), dtype=torch.float32, device=dev)
    nbytes = lib.mmb_gram_workspace_bytes(n, d, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for mode in modes:
        os.environ['MMB_TC_COLLECT'] = str(mode[0])
        os.environ['MMB_TC_SEG'] = str(mode[1] if len(mode) > 1 else 64)
        os.environ['MMB_TC_PASSES'] = str(mode[2] if len(mode) > 2 else 3)
        os.environ['MMB_TC_OPT'] = str(mode[3] if len(mode) > 3 else 15)
        ms = []
        for it in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nv.check(lib.mmb_gram(nv.ptr(X), n, d, nv.ptr(G), nv.ptr(ws), nbytes, 0, nv.stream_ptr()))
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        err = float((G.double() - want).abs().max() / want.abs().max())
        sym = bool(torch.equal(G, G.T))
        print(json.dumps({'collect': mode[0], 'seg': int(os.environ['MMB_TC_SEG']), 'passes': int(os.environ['MMB_TC_PASSES']), 'opt': int(os.environ['MMB_TC_OPT']), 'n': n, 'ms': float(np.mean(ms[3:])), 'best_ms': float(np.min(ms[3:])),
                          'tflops_algorithmic': 2.0 * n * d * d / (float(np.mean(ms[3:])) * 1e-3) / 1e12,
                          'max_rel_err_vs_f64': err, 'symmetric': sym}), flush=True)


if __name__ == '__main__':
    main()
