"""Host <-> device copy bandwidth per visible GPU (pinned memory), one at a time and all at once, plus the
NUMA facts that explain it.  Diagnostic for the e2e numbers of bench.py (which are PCIe / host-memory bound)."""
import glob, os, subprocess, sys, threading, time
import torch

def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return 'n/a (%s)' % e

print('cpus allowed:', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], '...')
print('cpuset:', sh('cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null || cat /sys/fs/cgroup/cpuset/cpuset.cpus'))
print('mems:', sh('cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null || cat /sys/fs/cgroup/cpuset/cpuset.mems'))
print('numa nodes:', sh('ls -d /sys/devices/system/node/node* | wc -l'), sh('cat /sys/devices/system/node/node*/cpulist | tr "\\n" " "'))
print(sh('nvidia-smi topo -m | head -14'))
n = torch.cuda.device_count()
for i in range(n):
    bdf = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), 'pci_bus_id') else ''
    q = sh('nvidia-smi -i %d --query-gpu=pci.bus_id --format=csv,noheader' % i).lower()
    node = sh('cat /sys/bus/pci/devices/%s/numa_node' % q[4:] if q.startswith('0000') else 'cat /sys/bus/pci/devices/%s/numa_node' % q)
    print('gpu', i, q, 'numa_node', node)
GB = 1 << 30
bufs = []
for i in range(n):
    torch.cuda.set_device(i)
    h = torch.empty(GB, dtype=torch.uint8).pin_memory()
    d = torch.empty(GB, dtype=torch.uint8, device='cuda:%d' % i)
    bufs.append((h, d, torch.cuda.Stream(device=i)))

def run(i, direction, reps=3):
    h, d, s = bufs[i]
    torch.cuda.set_device(i)
    with torch.cuda.stream(s):
        for _ in range(reps):
            (d.copy_(h, non_blocking=True) if direction == 'h2d' else h.copy_(d, non_blocking=True))
    s.synchronize()

for direction in ('h2d', 'd2h'):
    for i in range(n):
        run(i, direction, 1)
        t0 = time.perf_counter(); run(i, direction); dt = time.perf_counter() - t0
        print('%s gpu %d alone: %.1f GB/s' % (direction, i, 3 * GB / dt / 1e9))
    if n > 1:
        th = [threading.Thread(target=run, args=(i, direction)) for i in range(n)]
        t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt = time.perf_counter() - t0
        print('%s all %d GPUs at once: %.1f GB/s aggregate' % (direction, n, n * 3 * GB / dt / 1e9))
