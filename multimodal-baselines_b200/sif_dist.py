"""Utterance-sharded SIF embedding over the GPUs of one box.

The reference has no distributed code; this is the one data-parallel strategy the path
admits.  Utterances are split into contiguous blocks, one per rank (one process per GPU);
the table and the vocabulary weights are replicated.  Per split of the data:

    rank-local:  emb_r = weighted average of its block              (no communication)
                 G_r   = emb_r^T emb_r                               (d x d float32)
    exchange:    G = sum_r G_r            ONE all-reduce of d*d floats (360 KB at d = 300)
                 (N < d only: the start block S0 = X^T Omega is summed the same way)
    replicated:  components from (G, seeded start block) -- same bits on every rank, so no
                 broadcast is needed
    rank-local:  emb_r -= (emb_r pc^T) pc

Works with any torch.distributed backend: NCCL over NVLink on the GPUs; the partition /
reduction arithmetic is also exercised with gloo on CPU in tests/test_dist_cpu.py.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_global, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank`; the first n_global % world_size ranks get
    one extra row (same rule as numpy.array_split)."""
    base, rem = divmod(int(n_global), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t, group=None):
    """In-place sum over ranks (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def local_start_block(emb_local, n_global, lo, npc, start_block_fn):
    """N_global < d: this rank's share of S0 = X^T Omega, using rows [lo, hi) of the global
    seeded Omega (float64, d x (npc+10)); the caller all-reduces it."""
    omega = torch.as_tensor(start_block_fn(n_global, npc)[lo:lo + emb_local.shape[0]]).to(emb_local.device)
    return emb_local.double().T @ omega


def sharded_sif_embedding(table_t, vocab_w_t, ids_local_t, n_global, lo, npc=1, group=None, gram_mode=0,
                          timers=None):
    """SIF embedding + PC removal of this rank's block of a split of `n_global` utterances.

    table_t (V, d) f32, vocab_w_t (V,) f32, ids_local_t (n_local, L) int64 -- CUDA tensors on
    this rank's device.  Returns (emb_local (n_local, d) f32, pc (npc, d) f32).
    `timers`, if given, is a callable mark(name) invoked between stages (bench.py records
    CUDA events there)."""
    import _native as nv
    import sif_functions as sf
    from _native import lib

    mark = timers or (lambda name: None)
    n_local, L = ids_local_t.shape
    V, d = table_t.shape
    dev = table_t.device
    emb = torch.empty((n_local, d), dtype=torch.float32, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    nv.check(lib.mmb_sif_embed(nv.ptr(table_t), V, d, nv.ptr(vocab_w_t), nv.ptr(ids_local_t), n_local, L,
                               nv.ptr(emb), nv.ptr(st), nv.stream_ptr()))
    mark('embed')
    if npc <= 0:
        return emb, None, st
    if n_local > 0:
        G = sf.gram(emb, gram_mode)
    else:
        G = torch.zeros((d, d), dtype=torch.float32, device=dev)
    mark('gram')
    allreduce_sum_(G, group)
    S0 = None
    if n_global < d:
        S0 = local_start_block(emb, n_global, lo, npc, sf.start_block)
        allreduce_sum_(S0, group)
    mark('allreduce')
    pc = sf.pc_from_gram(G, npc, n_global, S0_t=S0)
    mark('pc')
    if n_local > 0:
        sf.project_out(emb, pc, out=emb)
    mark('project')
    return emb, pc, st


def cpu_reference_partition(X_blocks):
    """Host-side statement of the exchange step used by the CPU tests: the sum of per-shard
    Grams equals the Gram of the concatenation."""
    return sum(np.asarray(b, dtype=np.float64).T @ np.asarray(b, dtype=np.float64) for b in X_blocks)
