#!/usr/bin/env python
"""bench.py -- utterances/sec of the SIF hot path (embed + PC removal) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[3], SURVEY.md §8d "Config 4", the shape the roofline is
quoted on; it fits one GPU): N = 10,000,000 utterances x 64 tokens, 400,000-word vocabulary,
d = 300, Zipf(1.1) ids, lengths ~U[16,64] right-padded with id 0, synthetic GloVe-like
table with a planted common direction.  One step = one pass of the whole batch through
get_sentence_embeddings: gather + weighted average -> Gram -> principal component ->
projection subtraction.  With N > 1 ranks the 10 M utterances are sharded (strong scaling)
and the 300x300 Gram is summed with one NCCL all-reduce.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs in HBM);
`e2e` = the same metric through the host-buffer API (pinned ids in, embeddings out, copies
inside the timed region); `roofline` = the dominant kernel (the gather/average) against the
measured HBM copy bandwidth; `cpu_baseline` = the oracle port of the reference's NumPy /
sklearn path timed on this box's host cores (bounded sample).

`--impl reference` times that CPU port instead (rank 0 only) and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'multimodal-baselines_b200')
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

OUT = sys.stdout
METRIC = 'utterances/sec (SIF embed + PC removal)'
UNIT = 'utterances/s'
N_UTT = int(os.environ.get('MMB_BENCH_N', 10_000_000))
L_TOK = int(os.environ.get('MMB_BENCH_L', 64))
VOCAB = int(os.environ.get('MMB_BENCH_V', 400_000))
DIM = 300
ZIPF_S = 1.1
# algorithmic bytes per utterance of the embed pass at L tokens, d floats:
# ids L*8 + gathered rows L*d*4 (every token counted, no cache credit) + output d*4.
EMBED_BYTES_PER_UTT = L_TOK * 8 + L_TOK * DIM * 4 + DIM * 4
# kernels of libmmb_b200.so per step: embed 1, Gram 2 (tcgen05 + partial reduce), component solve
# n_iter + 3 = 10 (prep, 8 iterations, final), projection 1
LAUNCHES_PER_STEP = 14
WORKLOAD = 'sif_%dM_utt_x%d_tok_v%dk_d%d' % (N_UTT // 1_000_000, L_TOK, VOCAB // 1000, DIM) \
    if N_UTT >= 1_000_000 else 'sif_%d_utt_x%d_tok_v%d_d%d' % (N_UTT, L_TOK, VOCAB, DIM)


# ----------------------------------------------------------------------------- inputs
def zipf_pmf(V, s=ZIPF_S):
    p = 1.0 / np.arange(1, V, dtype=np.float64) ** s
    return p / p.sum()


def make_table_and_weights(device, V=VOCAB, d=DIM, seed=0):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    table = 0.4 * torch.randn((V, d), device=device, generator=g)
    table += 0.3 * torch.randn((1, d), device=device, generator=g)      # planted common direction
    table[0] = 0.0
    p = zipf_pmf(V)
    w = np.empty(V, dtype=np.float64)
    w[0] = 1.0
    w[1:] = 1e-3 / (1e-3 + p)                                            # sif.py:14-32 form
    return table.contiguous(), torch.as_tensor(w.astype(np.float32)).to(device), p


IDS_KIND = os.environ.get('MMB_BENCH_IDS', 'zipf')     # 'uniform': cache-hostile ids (secondary measurement)


def make_ids(device, n_rows, L, p, seed, out=None, block=1 << 20):
    """Zipf ids over 1..V-1 by inverse-CDF on the device, right-padded with 0."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if IDS_KIND == 'uniform':
        p = np.full(p.size, 1.0 / p.size)
    cdf = torch.as_tensor(np.cumsum(p)).to(device=device, dtype=torch.float32)
    ids = out if out is not None else torch.empty((n_rows, L), dtype=torch.int64, device=device)
    ar = torch.arange(L, device=device)[None, :]
    for s in range(0, n_rows, block):
        e = min(n_rows, s + block)
        u = torch.rand((e - s, L), device=device, generator=g)
        tok = torch.searchsorted(cdf, u).clamp_(max=p.size - 1) + 1
        lens = torch.randint(16 if L >= 16 else 1, L + 1, (e - s, 1), device=device, generator=g)
        tok[ar >= lens] = 0
        ids[s:e] = tok
    return ids


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def window(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        rows = [l for t, l in self.lines if t0 - 0.15 <= t <= t1 + 0.15] or [l for _, l in self.lines[-3:]]
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, val in zip(names, f[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()


# ----------------------------------------------------------------------------- CPU arm
def cpu_port_rate(table_np, weights_np, ids_np, repeats=1):
    """The oracle port of the reference path (its own Python loops + sklearn) on host cores."""
    from oracle import sif_oracle as so
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        so.get_sentence_embeddings_loop(table_np, weights_np, ids_np)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return ids_np.shape[0] / best, best


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from oracle import sif_oracle as so
    sample = int(os.environ.get('MMB_REF_SAMPLE', 20_000))
    rng = np.random.default_rng(0)
    p = zipf_pmf(VOCAB)
    table = (0.4 * rng.standard_normal((VOCAB, DIM), dtype=np.float32)
             + 0.3 * rng.standard_normal((1, DIM), dtype=np.float32))
    table[0] = 0
    weights = np.concatenate([[1.0], 1e-3 / (1e-3 + p)])
    cdf = np.cumsum(p)
    ids = np.minimum(np.searchsorted(cdf, rng.random((sample, L_TOK))), p.size - 1) + 1
    lens = rng.integers(16, L_TOK + 1, size=(sample, 1))
    ids[np.arange(L_TOK)[None, :] >= lens] = 0
    ids = ids.astype(np.int64)
    for _ in range(args.warmup):
        so.get_sentence_embeddings_loop(table, weights, ids)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        so.get_sentence_embeddings_loop(table, weights, ids)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    cores = cpu_threads()
    desc = ('%d-utterance sample of the %s workload per step; oracle port of sif.py:84-94 with the '
            "reference's own Python loops + sklearn TruncatedSVD" % (sample, WORKLOAD))
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample_utterances_per_step': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }), file=OUT, flush=True)


# ----------------------------------------------------------------------------- GPU arm
def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md: 6.65 TB/s)'


def load_tensor_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['bf16_tflops'])
    except Exception:
        return 1590.0


def measure_tf32_peak(dev, n=8192, reps=10):
    """Dense TF32 GEMM rate of this GPU, measured the way MEASURED_PEAKS.json measures bf16
    (torch.matmul n^3, 2 n^3 FLOP, best of `reps`, CUDA events): the driver records no TF32 figure,
    and the Gram kernel's pipe is kind::tf32.  Runs after the timed regions."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn((n, n), device=dev)
        b = torch.randn((n, n), device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def load_traffic(n_local):
    """DRAM bytes per embed launch from the committed ncu capture, if it was taken at this size."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'embed_traffic.json')) as fh:
            t = json.load(fh)
        if int(t.get('utterances_per_launch', -1)) == int(n_local):
            return float(t['dram_bytes_per_launch'])
        return float(t['dram_bytes_per_utterance']) * n_local if 'dram_bytes_per_utterance' in t else None
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    # stdout carries exactly ONE JSON line: keep the real stdout for it and point fd 1 at stderr, so that
    # whatever else writes to fd 1 (NCCL's version banner at NCCL_DEBUG >= VERSION, library prints, the
    # reference helpers' progress lines) cannot land in front of it
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference_arm(args)
        return

    if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
        os.environ['NCCL_DEBUG'] = 'WARN'           # NCCL's version banner goes to stdout: keep it to the one JSON line
    import torch.distributed as dist
    import _native as nv
    import sif_functions as sf
    import sif_dist as mdist
    from _native import lib

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    warmup = max(args.warmup, 3)

    lo, hi = mdist.shard_bounds(N_UTT, world, rank)
    n_local = hi - lo
    table, vocab_w, p = make_table_and_weights(dev)
    ids = make_ids(dev, n_local, L_TOK, p, seed=1000 + rank)
    torch.cuda.synchronize()

    # ---- device-resident steps, per-stage CUDA events on the launching stream ------------
    stages = ('embed', 'gram', 'allreduce', 'pc', 'project')
    ev = {}

    def run_step(record=None):
        marks = []

        def mark(name):
            if record is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))
        out = mdist.sharded_sif_embedding(table, vocab_w, ids, N_UTT, lo, npc=1, timers=mark)
        if record is not None:
            record.append(marks)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        emb, pc, st = run_step()
    barrier()
    nv.raise_on_status(st, VOCAB)
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    records = []
    start = torch.cuda.Event(enable_timing=True)
    stop = torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    start.record()
    step_starts = []
    for _ in range(args.steps):
        s = torch.cuda.Event(enable_timing=True)
        s.record()
        step_starts.append(s)
        emb, pc, st = run_step(records)
    stop.record()
    barrier()
    t_wall1 = time.time()
    ms_total = start.elapsed_time(stop)
    stage_ms = {k: 0.0 for k in stages}
    for s0, marks in zip(step_starts, records):
        prev = s0
        for name, e in marks:
            stage_ms[name] += prev.elapsed_time(e)
            prev = e
    stage_ms = {k: v / args.steps for k, v in stage_ms.items()}
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = N_UTT / (ms_per_step * 1e-3)
    clocks = sampler.window(t_wall0, t_wall1)

    # ---- end to end: pinned host ids in, float32 embeddings out, copies inside the region --
    e2e = None
    if not args.no_e2e:
        h_ids = nv.PinnedArray((n_local, L_TOK), np.int64)
        h_out = nv.PinnedArray((n_local, DIM), np.float32)
        torch.as_tensor(h_ids.array).copy_(ids)          # D2H once, outside the timed region
        torch.cuda.synchronize()
        omega = np.ascontiguousarray(sf.start_block(DIM if N_UTT >= DIM else N_UTT, 1))

        def e2e_step():
            if world == 1:
                nv.check(lib.mmb_sif_embedding_host(nv.ptr(table), VOCAB, DIM, nv.ptr(vocab_w),
                                                    nv.np_ptr(h_ids.array), n_local, L_TOK, 1, nv.np_ptr(omega),
                                                    nv.np_ptr(h_out.array), 0, None, nv.GRAM_AUTO, 0))
            elif mdist.default_comm() is not None and N_UTT >= DIM:
                mdist.default_comm().sif_embedding_host(table, vocab_w, h_ids.array, h_out.array, N_UTT, npc=1)
            else:
                d_ids = torch.as_tensor(h_ids.array).to(dev, non_blocking=True)
                e, _pc, _st = mdist.sharded_sif_embedding(table, vocab_w, d_ids, N_UTT, lo, npc=1)
                torch.as_tensor(h_out.array).copy_(e, non_blocking=True)
                torch.cuda.synchronize()
        del emb
        torch.cuda.empty_cache()
        for _ in range(2):
            e2e_step()
        barrier()
        n_e2e = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {'value': N_UTT / float(dt.item()), 'unit': UNIT,
               'h2d_bytes_per_step': int(N_UTT) * L_TOK * 8, 'd2h_bytes_per_step': int(N_UTT) * DIM * 4,
               'ms_per_step': float(dt.item()) * 1e3, 'steps': n_e2e,
               'api': 'mmb_sif_embedding_host (C ABI, pinned host buffers, float32 out)' if world == 1 else
                      ('mmb_sif_embedding_host_peer (C ABI, pinned host buffers per rank, Gram summed over NVLink)'
                       if mdist.default_comm() is not None else
                       'pinned ids -> dist.sharded_sif_embedding -> pinned float32 out, per rank')}
        h_ids.free()
        h_out.free()

    # ---- roofline of the dominant kernel + CPU baseline (rank 0) ---------------------------
    peak, peak_src = load_peaks()
    tf32_peak = measure_tf32_peak(dev) if rank == 0 else None
    tf32_src = 'measured here: torch.matmul TF32 8192^3, best of 10'
    if not tf32_peak:
        tf32_peak, tf32_src = load_tensor_peak() / 2.0, 'bf16_tflops / 2 (TF32, assumed)'
    embed_s = stage_ms['embed'] * 1e-3
    traffic = load_traffic(n_local) if IDS_KIND == 'zipf' else None
    achieved = n_local * EMBED_BYTES_PER_UTT / embed_s / 1e9
    roofline = {'bound': 'hbm', 'kernel': 'sif_embed_warp_prefetch_kernel<3,2,4>', 'achieved': achieved, 'peak': peak,
                'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                # what HBM itself carried (ncu DRAM bytes of the committed capture / this run's launch time):
                # the gap to `achieved` is rows served by L1/L2 (Zipf head + merged pad runs)
                'dram_gbs': (traffic / embed_s / 1e9) if traffic else None,
                'dram_frac': (traffic / embed_s / 1e9 / peak) if traffic else None,
                'peak_source': peak_src, 'algorithmic_bytes_per_launch': n_local * EMBED_BYTES_PER_UTT,
                'launch_ms': stage_ms['embed'], 'stage_ms': stage_ms,
                # the other two streaming stages against their own bounds: the Gram's
                # algorithmic 2 d^2 FLOP per utterance against half the measured bf16 GEMM rate (no TF32
                # rate is measured on this pool; the 3xTF32 kernel executes 2.05x the algorithmic FLOPs),
                # the projection's 2 d 4 bytes per utterance against the measured copy bandwidth
                'other_stages': {
                    'gram': {'bound': 'tensor', 'unit': 'TFLOP/s',
                             'achieved': n_local * 2.0 * DIM * DIM / (stage_ms['gram'] * 1e-3) / 1e12,
                             'peak': tf32_peak, 'peak_source': tf32_src, 'executed_over_algorithmic': 2.05},
                    'project': {'bound': 'hbm', 'unit': 'GB/s',
                                'achieved': n_local * 2.0 * DIM * 4 / (stage_ms['project'] * 1e-3) / 1e9,
                                'peak': peak}}}
    for _st in roofline['other_stages'].values():
        _st['frac'] = _st['achieved'] / _st['peak']
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = int(os.environ.get('MMB_CPU_SAMPLE', 50_000))
        sample = min(sample, n_local)
        rate, secs = cpu_port_rate(table.cpu().numpy(), vocab_w.double().cpu().numpy(), ids[:sample].cpu().numpy())
        cpu = {'value': rate, 'unit': UNIT, 'cores': cpu_threads(), 'kind': 'port',
               'sample': 'first %d utterances of this workload, %.1f s; oracle port of sif.py:84-94 with the '
                         "reference's Python loops + sklearn TruncatedSVD (BLAS threads = all cores)" % (sample, secs)}
    sampler.stop()

    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'utterances': N_UTT, 'tokens_per_utterance': L_TOK, 'vocab': VOCAB,
                       'dim': DIM, 'ids': ('uniform over the vocabulary (cache-hostile variant)' if IDS_KIND == 'uniform'
                                           else 'Zipf(1.1)') + ', lengths U[16,64], pad id 0', 'npc': 1,
                       'parallelism': 'utterance shards x%d + 1 all-reduce of the 300x300 Gram (%s)' % (
                           world, 'NVLink peer memory, fused into the Gram reduce kernel' if (world > 1 and mdist.default_comm() is not None) else 'NCCL' if world > 1 else 'none at 1 GPU'),
                       'l2': 'inputs larger than L2 (ids %.1f GB + embeddings %.1f GB per rank, table 0.48 GB)'
                             % (n_local * L_TOK * 8 / 1e9, n_local * DIM * 4 / 1e9)},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': LAUNCHES_PER_STEP * args.steps,
            'roofline': roofline, 'cpu_baseline': cpu,
        }), file=OUT, flush=True)
    if world > 1:
        mdist.close_default_comms()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
