"""Time the embed kernel alone on a slice of the bench workload (variant / grid via env)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import torch, bench, _native as nv
from _native import lib
n = int(os.environ.get('PROFILE_N', 4_000_000))
dev = torch.device('cuda')
table, vw, p = bench.make_table_and_weights(dev)
ids = bench.make_ids(dev, n, bench.L_TOK, p, seed=1000)
if os.environ.get('NOPAD'):
    ids2 = bench.make_ids(dev, n, bench.L_TOK, p, seed=7)
    ids = torch.where(ids == 0, ids2.clamp(min=1), ids)
emb = torch.empty((n, 300), device=dev)
st = torch.zeros(1, dtype=torch.int32, device=dev)
def run():
    nv.check(lib.mmb_sif_embed(nv.ptr(table), table.shape[0], 300, nv.ptr(vw), nv.ptr(ids), n, bench.L_TOK, nv.ptr(emb), nv.ptr(st), nv.stream_ptr()))
for _ in range(3): run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): run()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print('variant %s waves %s nopad %s: %.3f ms for %d utt -> %.1f M utt/s, %.0f GB/s algorithmic' % (
    os.environ.get('MMB_EMBED_VARIANT', '0'), os.environ.get('MMB_EMBED_WAVES', '8'), os.environ.get('NOPAD', '0'),
    ms, n, n / ms / 1e3, n * bench.EMBED_BYTES_PER_UTT / ms / 1e6))
