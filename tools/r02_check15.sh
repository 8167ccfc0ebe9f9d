set -x
mkdir -p gpurun_out
SECONDS=0
python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$? wall=${SECONDS}s"; python - <<PY
import json
d=json.load(open('gpurun_out/r02c_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(json.dumps(d['secondary'])[:900]); print(json.dumps(d['cpu_baseline'])[:400]); print(d['verify']['ok'], json.dumps(d['roofline'])[:700])
PY
tail -3 gpurun_out/r02c_bench_n1.err | cut -c1-300
SECONDS=0
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err; echo "ref rc=$? wall=${SECONDS}s"; cut -c1-300 gpurun_out/r02c_bench_ref.json
