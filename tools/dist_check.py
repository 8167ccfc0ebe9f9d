"""torchrun --nproc-per-node R tools/dist_check.py : the utterance-sharded SIF path over NCCL
must reproduce the single-GPU result (same PC up to FP32 reduction order, same embeddings)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import numpy as np
import torch
import torch.distributed as dist
import sif_dist
import sif_functions as sf


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    ok = True
    for n_global, L, V in ((100_003, 64, 50_000), (157, 20, 3016)):      # N >= d and N < d (transposed) cases
        g = torch.Generator(device=dev)
        g.manual_seed(5)                                                 # same data on every rank
        table = 0.4 * torch.randn((V, 300), device=dev, generator=g) + 0.3 * torch.randn((1, 300), device=dev, generator=g)
        table[0] = 0
        vw = torch.rand(V, device=dev, generator=g) * 0.9 + 0.1
        ids = torch.randint(1, V, (n_global, L), device=dev, generator=g)
        lens = torch.randint(1, L + 1, (n_global, 1), device=dev, generator=g)
        ids[torch.arange(L, device=dev)[None, :] >= lens] = 0
        lo, hi = sif_dist.shard_bounds(n_global, world, rank)
        emb_l, pc_l, st = sif_dist.sharded_sif_embedding(table, vw, ids[lo:hi].contiguous(), n_global, lo, npc=1)
        # the NVLink peer exchange (default) against the NCCL all-reduce
        emb_n, pc_n, _ = sif_dist.sharded_sif_embedding(table, vw, ids[lo:hi].contiguous(), n_global, lo, npc=1,
                                                        comm=None)
        comm = sif_dist.default_comm()
        assert comm is not None, 'PeerComm was not created'
        comm.check()
        cos_n = float((pc_l[0].double() @ pc_n[0].double()))
        err_n = float((emb_l - emb_n).abs().max() / emb_n.abs().max())
        # stand-alone peer all-reduce == rank-ordered sum, bit for bit
        x = torch.randn(90000, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
        xs = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(xs, x)
        want = xs[0].clone()
        for r in range(1, world):
            want += xs[r]
        got = comm.allreduce_(x.clone())
        comm.check()
        exact = torch.equal(got, want)
        # the vector of the data-parallel MMB step: 843,400 head-parameter gradients (3.4 MB, vector path + tail)
        for n_big, dt in ((843_401, torch.float32), (3_300, torch.float64)):
            xb = torch.randn(n_big, device=dev, dtype=dt, generator=torch.Generator(device=dev).manual_seed(7 + rank))
            xs = [torch.empty_like(xb) for _ in range(world)]
            dist.all_gather(xs, xb)
            want = xs[0].clone()
            for r in range(1, world):
                want += xs[r]
            got = comm.allreduce_(xb.clone())
            comm.check()
            exact = exact and torch.equal(got, want)
        if n_global >= 300:
            h_ids = ids[lo:hi].cpu().numpy()
            h_out = np.empty((hi - lo, 300), dtype=np.float32)
            comm.sif_embedding_host(table, vw, h_ids, h_out, n_global, npc=1, chunk_rows=7000)
            err_h = float(np.abs(h_out - emb_l.cpu().numpy()).max() / np.abs(h_out).max())
            print('rank %d N=%d: host-buffer peer path vs device path: emb err %.2e' % (rank, n_global, err_h), flush=True)
            ok = ok and err_h < 2e-5      # Gram summed per chunk: FP32 order differs from the one-shot Gram
        print('rank %d N=%d: peer vs NCCL: pc cos %.8f, emb err %.2e; peer all-reduce exact %s'
              % (rank, n_global, cos_n, err_n, exact), flush=True)
        ok = ok and cos_n > 0.9999999 and err_n < 1e-6 and exact
        emb_1, pc_1 = sf.sif_embedding_device(table, vw, ids, npc=1, return_pc=True)
        cos = float((pc_l[0].double() @ pc_1[0].double()))
        err = float((emb_l - emb_1[lo:hi]).abs().max() / emb_1.abs().max())
        pcs = [torch.empty_like(pc_l) for _ in range(world)]
        dist.all_gather(pcs, pc_l)
        same = all(torch.equal(pcs[0], p) for p in pcs)                  # replicated solve: identical bits
        good = cos > 0.99999 and err < 2e-5 and same
        ok = ok and good
        print('rank %d N=%d: pc cos %.8f, emb err %.2e, replicas identical %s -> %s'
              % (rank, n_global, cos, err, same, 'ok' if good else 'FAIL'), flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    sif_dist.close_default_comms()
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == '__main__':
    main()
