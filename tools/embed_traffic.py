"""profiles/embed_traffic.json from the ncu summaries of the embed kernel (tools/ncu_summary.py output):

    python tools/embed_traffic.py zipf=profiles/r02_ncu_kernels.json uniform=profiles/r02_ncu_embed_uniform_ids.json

One record per id distribution: the demangled kernel name ncu saw (bench.py refuses the record when the
library reports a different instantiation), utterances per profiled launch, DRAM bytes per utterance,
L1 LSU wavefront utilisation, L2->L1 bytes and their share of the measured LTS cap (6300 B/cycle full chip,
B300_MICROARCH.md "L2 cache")."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LTS_CAP_BYTES_PER_CYCLE = 6300.0


def num(s):
    return float(str(s).split()[0])


def to_bytes(s):
    v, u = str(s).split()[:2]
    return float(v) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}[u]


def to_ms(s):
    v, u = str(s).split()[:2]
    return float(v) * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}[u]


def record(path, n_utt):
    rows = [r for r in json.load(open(path)) if 'sif_embed' in r['Kernel Name']]
    r = max(rows, key=lambda e: to_ms(e['gpu__time_duration.sum']))   # the working kernel (the stand-by one exits at once)
    ms = to_ms(r['gpu__time_duration.sum'])
    dram = to_bytes(r['dram__bytes_read.sum']) + to_bytes(r['dram__bytes_write.sum'])
    l2l1 = to_bytes(r['l1tex__m_xbar2l1tex_read_bytes.sum'])
    cycles = num(r['sm__cycles_elapsed.max'])
    return {
        'ncu_kernel_name': r['Kernel Name'],
        'utterances_per_launch': n_utt,
        'duration_ms_under_ncu': ms,
        'dram_bytes_per_launch': dram,
        'dram_bytes_per_utterance': dram / n_utt,
        'dram_pct_of_ncu_peak': num(r['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']),
        'l1_hit_pct': num(r['l1tex__t_sector_hit_rate.pct']),
        'l2_hit_pct': num(r['lts__t_sector_hit_rate.pct']),
        'l1_lsu_wavefronts_pct': num(r['l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']),
        'l2_to_l1_bytes_per_utterance': l2l1 / n_utt,
        'l2_to_l1_pct_of_lts_cap': 100.0 * l2l1 / (cycles * LTS_CAP_BYTES_PER_CYCLE),
        'registers_per_thread': num(r['launch__registers_per_thread']),
        'source': '%s (ncu --set full, tools/profile_kernels.py, PROFILE_N=%d)' % (os.path.relpath(path, ROOT), n_utt),
    }


def main(argv):
    n_utt = int(os.environ.get('PROFILE_N', 2_000_000))
    out = {}
    for a in argv:
        kind, path = a.split('=', 1)
        out[kind] = record(path, n_utt)
    dst = os.path.join(ROOT, 'profiles', 'embed_traffic.json')
    old = {}
    if os.path.exists(dst):
        try:
            old = json.load(open(dst))
        except Exception:
            old = {}
    old = {k: v for k, v in old.items() if isinstance(v, dict)}
    old.update(out)
    json.dump(old, open(dst, 'w'), indent=1)
    print(json.dumps(old, indent=1))


if __name__ == '__main__':
    main(sys.argv[1:])
