"""Bring-up check of the tcgen05 3xTF32 Gram against float64 NumPy and the FP32 kernel.
Run on the GPU box: python tools/gram_tc_check.py [N ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import numpy as np
import torch
import _native as nv
import sif_functions as sf


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [16, 50, 300, 5000, 200_000]
    torch.manual_seed(0)
    for n in sizes:
        X = (0.4 * torch.randn(n, 300, device='cuda') + 0.3 * torch.randn(1, 300, device='cuda')).contiguous()
        want = (X.double().T @ X.double())
        G32 = sf.gram(X, nv.GRAM_FP32)
        t0 = time.perf_counter()
        Gtc = sf.gram(X, nv.GRAM_TF32X3)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        scale = want.abs().max().item()
        e32 = (G32.double() - want).abs().max().item() / scale
        etc = (Gtc.double() - want).abs().max().item() / scale
        sym = (Gtc - Gtc.T).abs().max().item()
        # where is the error? per 32x32 block max error map (coarse)
        err = ((Gtc.double() - want).abs() / scale).cpu().numpy()
        blocks = [[err[i:i + 64, j:j + 64].max() for j in range(0, 300, 64)] for i in range(0, 300, 64)]
        print('N=%d  fp32 err %.2e  tc err %.2e  sym %.1e  (%.1f ms incl. launch)' % (n, e32, etc, sym, dt * 1e3))
        if etc > 1e-4:
            print(np.array2string(np.array(blocks), precision=1, floatmode='maxprec'))
    # timing at bench size
    n = 2_000_000
    X = torch.randn(n, 300, device='cuda')
    for mode, name in ((nv.GRAM_TF32X3, 'tf32x3'), (nv.GRAM_FP32, 'fp32')):
        for _ in range(2):
            sf.gram(X, mode)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            sf.gram(X, mode)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 5
        print('gram %s N=%d: %.3f ms  -> %.1f TFLOP/s algorithmic (2*d^2 per row)' % (name, n, ms, n * 180000 / ms / 1e9))


if __name__ == '__main__':
    main()
