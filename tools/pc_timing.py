"""Phase clock log of the grid-parallel component solve (sets MMB_PC_TIMING=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
os.environ['MMB_PC_TIMING'] = '1'
import numpy as np, torch
import _native as nv, sif_functions as sf
from _native import lib
dev = torch.device('cuda')
torch.manual_seed(0)
X = (0.4 * torch.randn(200000, 300, device=dev) + 0.3 * torch.randn(1, 300, device=dev)).contiguous()
G = sf.gram(X, nv.GRAM_FP32)
for npc in (1, 3):
    k = npc + 10
    S0 = torch.as_tensor(sf.start_block(300, npc)).to(dev)
    pc = torch.empty((npc, 300), device=dev)
    nbytes = lib.mmb_pc_workspace_bytes(300, k)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        nv.check(lib.mmb_pc_from_gram(nv.ptr(G), 300, nv.ptr(S0), k, npc, 0, 7, nv.ptr(pc), nv.ptr(ws), nbytes, nv.stream_ptr()))
    torch.cuda.synchronize()
    t = ws[nbytes - 4096:].view(torch.int64).cpu().numpy().reshape(-1, 16)
    print('npc', npc)
    print('  iter1 (load+sum | chol | apply | mult | reduce+partials):', np.diff(t[1][:6]).tolist())
    print('  iter7 (2 passes):', np.diff(t[7][:6]).tolist())
    print('  final (Tsum | chol+apply | WtW | eig | components):', np.diff(t[10][:6]).tolist())
    print('  kernel start-to-start clocks:', np.diff(t[:8, 0]).tolist(), ' final-iter7 start', t[10][0] - t[7][0])
