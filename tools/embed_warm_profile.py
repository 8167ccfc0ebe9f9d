"""ncu driver: the pre-scaled embed kernel with the warm-row L1 policy (K = MMB_EMBED_WARM) on a 2M-utterance slice."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import torch, bench
import _native as nv
from _native import lib
dev = torch.device('cuda')
table, vw, p = bench.make_table_and_weights(dev)
ids = bench.make_ids(dev, 2_000_000, bench.L_TOK, p, seed=1000)
n, L = ids.shape; V, d = table.shape
emb = torch.empty((n, d), dtype=torch.float32, device=dev); st = torch.zeros(1, dtype=torch.int32, device=dev)
nbytes = lib.mmb_sif_embed_workspace_bytes(V, d, n, L); ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
for _ in range(2):
    nv.check(lib.mmb_sif_embed_ws(nv.ptr(table), V, d, nv.ptr(vw), nv.ptr(ids), n, L, nv.ptr(emb), nv.ptr(st), nv.ptr(ws), nbytes, nv.stream_ptr()))
torch.cuda.synchronize(); print('ok', (lib.mmb_last_kernel(0) or b'').decode())
