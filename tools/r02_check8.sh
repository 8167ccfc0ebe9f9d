set -x
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_gpus_n$N.txt 2>&1
nproc >> gpurun_out/r02_gpus_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/dist_check.py > gpurun_out/r02_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep -E "replicas|peer vs|host-buffer" gpurun_out/r02_dist_check_n$N.log | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 tools/dp_check.py > gpurun_out/r02_dp_check_n$N.log 2>&1; echo "dp_check rc=$?"; grep -E "^rank 0|ms_per_step|Error" gpurun_out/r02_dp_check_n$N.log | tail -8
timeout 300 python tools/pcie_probe.py > gpurun_out/r02_pcie_probe_n$N.log 2>&1; tail -22 gpurun_out/r02_pcie_probe_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(json.dumps(d['e2e'])[:1200]); print(json.dumps(d['verify'])[:1500]); print(d['roofline']['stage_ms'])
PY
tail -3 gpurun_out/r02_bench_n$N.err
P=$((N*${2:-4}))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $P --master-addr 127.0.0.1 --master-port 29554 multimodal-baselines_b200/sweep.py --limit ${3:-512} --out gpurun_out/r02_sweep_full_n$N.jsonl > gpurun_out/r02_sweep_full_n$N.json 2> gpurun_out/r02_sweep_full_n$N.err; echo "sweep rc=$?"; tail -1 gpurun_out/r02_sweep_full_n$N.json | cut -c1-900
