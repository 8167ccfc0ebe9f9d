"""Host <-> device copy bandwidth of every visible GPU (pinned memory): one at a time, all at once, both
directions at once, plain vs write-combined host buffers and a chunk-size sweep, plus the NUMA facts that explain
the numbers.  Diagnostic for bench.py's `e2e` (PCIe / host-memory bound): what can this box move at best?

    python tools/pcie_probe.py [--gb 1]        # one process, one thread per GPU
"""
import argparse
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'multimodal-baselines_b200'))
import numpy as np
import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return 'n/a (%s)' % e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gb', type=float, default=1.0)
    a = ap.parse_args()
    import _native as nv
    print('cpus allowed:', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], '...')
    print('cpuset:', sh('cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null || cat /sys/fs/cgroup/cpuset/cpuset.cpus'))
    print('mems:', sh('cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null || cat /sys/fs/cgroup/cpuset/cpuset.mems'))
    print('numa nodes:', sh('ls -d /sys/devices/system/node/node* | wc -l'),
          sh('cat /sys/devices/system/node/node*/cpulist | tr "\\n" " "'))
    print('host memory:', sh("grep -E 'MemTotal|MemAvailable' /proc/meminfo | tr '\\n' ' '"))
    print(sh('nvidia-smi topo -m | head -14'))
    n = torch.cuda.device_count()
    for i in range(n):
        prop = torch.cuda.get_device_properties(i)
        bdf = '%04x:%02x:%02x.0' % (getattr(prop, 'pci_domain_id', 0), prop.pci_bus_id, prop.pci_device_id)
        print('gpu', i, bdf, 'numa_node', sh('cat /sys/bus/pci/devices/%s/numa_node' % bdf),
              'link', sh('cat /sys/bus/pci/devices/%s/current_link_speed /sys/bus/pci/devices/%s/current_link_width' % (bdf, bdf)).replace('\n', ' x'))
    nbytes = int(a.gb * (1 << 30))
    bufs = []
    for i in range(n):
        torch.cuda.set_device(i)
        h = nv.PinnedArray((nbytes,), np.uint8)
        hw = nv.PinnedArray((nbytes,), np.uint8, write_combined=True)
        d = torch.empty(nbytes, dtype=torch.uint8, device='cuda:%d' % i)
        d2 = torch.empty(nbytes, dtype=torch.uint8, device='cuda:%d' % i)
        bufs.append(dict(h=torch.from_numpy(h.array), hw=torch.from_numpy(hw.array), d=d, d2=d2, keep=(h, hw),
                         s=torch.cuda.Stream(device=i), s2=torch.cuda.Stream(device=i)))

    def run(i, direction, reps=3, host='h', chunk=0):
        b = bufs[i]
        torch.cuda.set_device(i)
        h, d = b[host], b['d']
        with torch.cuda.stream(b['s']):
            for _ in range(reps):
                if chunk:
                    for o in range(0, nbytes, chunk):
                        (d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True) if direction == 'h2d'
                         else h[o:o + chunk].copy_(d[o:o + chunk], non_blocking=True))
                else:
                    (d.copy_(h, non_blocking=True) if direction == 'h2d' else h.copy_(d, non_blocking=True))
        if direction == 'both':
            pass
        b['s'].synchronize()

    def run_both(i, reps=3):
        b = bufs[i]
        torch.cuda.set_device(i)
        for _ in range(reps):
            with torch.cuda.stream(b['s']):
                b['d'].copy_(b['hw'], non_blocking=True)
            with torch.cuda.stream(b['s2']):
                b['h'].copy_(b['d2'], non_blocking=True)
        b['s'].synchronize()
        b['s2'].synchronize()

    def timed(fn, gpus, *args):
        th = [threading.Thread(target=fn, args=(i,) + args) for i in gpus]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        return time.perf_counter() - t0

    allg = list(range(n))
    for direction in ('h2d', 'd2h'):
        for i in allg:
            run(i, direction, 1)
            dt = timed(run, [i], direction)
            print('%s gpu %d alone: %.1f GB/s' % (direction, i, 3 * nbytes / dt / 1e9))
        if n > 1:
            for k in sorted(set([2, 4, n]) & set(range(2, n + 1))):
                dt = timed(run, allg[:k], direction)
                print('%s %d GPUs at once: %.1f GB/s aggregate (%.1f per GPU)' % (direction, k, k * 3 * nbytes / dt / 1e9, 3 * nbytes / dt / 1e9))
    dt = timed(run, allg, 'h2d', 3, 'hw')
    print('h2d from WRITE-COMBINED host memory, %d GPUs at once: %.1f GB/s aggregate' % (n, n * 3 * nbytes / dt / 1e9))
    dt = timed(run, [0], 'h2d', 3, 'hw')
    print('h2d from WRITE-COMBINED host memory, gpu 0 alone: %.1f GB/s' % (3 * nbytes / dt / 1e9))
    for chunk_mb in (4, 32, 128):
        dt = timed(run, allg, 'h2d', 3, 'h', chunk_mb << 20)
        print('h2d in %d MiB chunks, %d GPUs at once: %.1f GB/s aggregate' % (chunk_mb, n, n * 3 * nbytes / dt / 1e9))
        dt = timed(run, allg, 'd2h', 3, 'h', chunk_mb << 20)
        print('d2h in %d MiB chunks, %d GPUs at once: %.1f GB/s aggregate' % (chunk_mb, n, n * 3 * nbytes / dt / 1e9))
    dt = timed(run_both, [0])
    print('h2d + d2h at once, gpu 0 alone: %.1f GB/s each way' % (3 * nbytes / dt / 1e9))
    dt = timed(run_both, allg)
    print('h2d + d2h at once, %d GPUs: %.1f GB/s each way aggregate' % (n, n * 3 * nbytes / dt / 1e9))


if __name__ == '__main__':
    main()
