set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu14.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu14.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke14.log 2>&1; tail -3 gpurun_out/r02_smoke14.log
/usr/bin/time -v python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench rc=$?"; grep -E "Elapsed|Maximum resident" gpurun_out/r02c_bench_n1.err; python - <<PY
import json
d=json.load(open('gpurun_out/r02c_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(json.dumps(d['secondary'])[:900]); print(json.dumps(d['cpu_baseline'])[:400]); print(d['verify']['ok'], json.dumps(d['roofline'])[:600])
PY
/usr/bin/time -v python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err; grep -E "Elapsed" gpurun_out/r02c_bench_ref.err; cut -c1-300 gpurun_out/r02c_bench_ref.json
