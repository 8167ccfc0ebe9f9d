set -x
mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_n$N.log 2>&1; tail -8 gpurun_out/r02_pytest_gpu_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/dist_check.py > gpurun_out/r02_dist_check_n$N.log 2>&1; echo "dist_check rc=$?"; grep -E "replicas|peer vs|host-buffer" gpurun_out/r02_dist_check_n$N.log | tail -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 tools/dp_check.py > gpurun_out/r02_dp_check_n$N.log 2>&1; echo "dp_check rc=$?"; grep -E "^rank|ms_per_step|Error" gpurun_out/r02_dp_check_n$N.log | tail -14
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(json.dumps(d['e2e'])[:900]); print(json.dumps(d['verify'])[:1500])
PY
tail -3 gpurun_out/r02_bench_n$N.err
