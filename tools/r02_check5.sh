set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu5.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu5.log | cut -c1-250
python tools/bench_mmb.py --shape all --steps 100 --no-cpu > gpurun_out/r02_bench_mmb.jsonl 2> gpurun_out/bm.err; python - <<PY
import json
for l in open('gpurun_out/r02_bench_mmb.jsonl'):
    if l.startswith('{'):
        d=json.loads(l); print(d['shape'], {k:round(v['ms_per_step'],4) for k,v in d.items() if isinstance(v,dict) and 'ms_per_step' in v})
PY
python tools/profile_kernels.py > gpurun_out/plain_prof.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sif_embed|prescale|gram_tc|remove_pc" -c 8 -o gpurun_out/prof_r02b -f python tools/profile_kernels.py > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
MMB_BENCH_IDS=uniform PROFILE_ITERS=1 python tools/profile_kernels.py > gpurun_out/plain_prof_u.log 2>&1 && MMB_BENCH_IDS=uniform PROFILE_ITERS=1 timeout 400 ncu --set full --clock-control none -k regex:"sif_embed_prescaled" -c 1 -o gpurun_out/prof_r02b_uniform -f python tools/profile_kernels.py > gpurun_out/ncu_u.log 2>&1; tail -2 gpurun_out/ncu_u.log
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --no-verify > gpurun_out/plain_short.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --no-verify > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
