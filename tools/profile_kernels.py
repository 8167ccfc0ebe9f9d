"""Small driver for ncu: one launch of each SIF kernel at a 2M-utterance slice of the bench
workload (same distributions as bench.py; sif_embedding_device takes the pre-scaled-table embed path for
batches of this size, exactly as bench.py's timed step does)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import torch
import bench
import _native as nv
import sif_functions as sf

n = int(os.environ.get('PROFILE_N', 2_000_000))
dev = torch.device('cuda')
table, vw, p = bench.make_table_and_weights(dev)
ids = bench.make_ids(dev, n, bench.L_TOK, p, seed=1000)
torch.cuda.synchronize()
for it in range(int(os.environ.get('PROFILE_ITERS', 2))):
    emb, pc = sf.sif_embedding_device(table, vw, ids, npc=1, return_pc=True)
    torch.cuda.synchronize()
print('ok', emb.shape, float(emb.abs().max()), float(pc.norm()))
