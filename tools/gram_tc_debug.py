"""Layout bring-up of the tcgen05 Gram: structured inputs that expose operand permutations."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import numpy as np
import torch
import _native as nv
import sif_functions as sf

torch.set_printoptions(precision=3, linewidth=200, sci_mode=False)


def show(name, X):
    want = (X.double().T @ X.double()).float()
    G, ws = sf.gram(X, nv.GRAM_TF32X3, return_ws=True)
    torch.cuda.synchronize()
    if os.environ.get('MMB_TC_DEBUG'):
        f = ws.view(torch.float32)
        dbg = f[-1024:].cpu().numpy()
        print('  dbg box0 rows0-3 (raw smem floats 0..127):', dbg[:40], '...', dbg[96:104])
        print('  dbg box1:', dbg[256:264])
        print('  tmem base word:', dbg[512:516].view(np.uint32))
        part = f[:128 * 480].view(128, 480)
        print('  partial CTA0 nnz', (part != 0).sum().item(), 'row0[:8]', part[0, :8].cpu().numpy(), 'row1[:8]', part[1, :8].cpu().numpy())
    print('==', name, 'passes', os.environ.get('MMB_TC_PASSES', '3'), 'lbo', os.environ.get('MMB_TC_LBO'), 'sbo',
          os.environ.get('MMB_TC_SBO'))
    print('max|G|', G.abs().max().item(), 'max|want|', want.abs().max().item(), 'nnz G', (G != 0).sum().item(),
          'nnz want', (want != 0).sum().item(), 'err', (G - want).abs().max().item())
    for (r, c) in ((0, 0), (0, 32), (1, 0), (33, 64), (128, 128), (130, 200), (256, 256), (260, 290)):
        print('  G[%d,%d:%d] ' % (r, c, c + 6), G[r, c:c + 6].cpu().numpy(), ' want ', want[r, c:c + 6].cpu().numpy())
    nz = (G != 0).nonzero()
    if 0 < nz.shape[0] <= 40:
        for i, j in nz.cpu().numpy():
            print('   nz', i, j, G[i, j].item())


n = 16
# 1: single row, value = column index + 1  -> G[i][j] = (i+1)(j+1)
X = torch.zeros(n, 300, device='cuda')
X[0] = torch.arange(1, 301, device='cuda').float()
show('row0 = 1..300', X)
# 2: single row 5
X = torch.zeros(n, 300, device='cuda')
X[5] = torch.arange(1, 301, device='cuda').float()
show('row5 = 1..300', X)
# 3: two unit entries -> exactly G[2][7] = G[7][2] = 0, G[2][2] = 1, G[7][7]=1 ... use row 3: e2 + 2 e7
X = torch.zeros(n, 300, device='cuda')
X[3, 2] = 1.0
X[3, 7] = 2.0
X[9, 40] = 3.0
X[9, 2] = 1.0
show('sparse', X)
