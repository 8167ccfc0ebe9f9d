"""Summarise an .ncu-rep (ncu --set full) into a small JSON: one entry per profiled launch with
the counters the roofline discussion uses.   python tools/ncu_summary.py REPORT.ncu-rep OUT.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']


def main(rep, out):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt[txt.index('"ID"'):])))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    res = []
    for r in data:
        e = {'Kernel Name': r[col['Kernel Name']][:160]}
        for k in KEYS:
            if k in col:
                e[k] = '%s %s' % (r[col[k]], units[col[k]])
        res.append(e)
    json.dump(res, open(out, 'w'), indent=1)
    for e in res:
        print(e['Kernel Name'][:60], e.get('gpu__time_duration.sum'), 'dram%', e.get(KEYS[3]), 'tensor%', e.get(KEYS[14]),
              'L1wf%', e.get(KEYS[6]))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
