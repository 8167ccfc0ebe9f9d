set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu4.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu4.log | cut -c1-250
python bench.py --no-secondary > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/r02b_bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['roofline']['kernel'], d['roofline']['stage_ms']); print(json.dumps(d['verify'])[:1200]); print(d['e2e']['value'], d['e2e']['copy_ceiling']['frac'])
print(json.dumps(d['roofline_hbm_bound'])[:600])
PY
for v in 1 2 3; do MMB_EMBED_PS_VARIANT=$v python bench.py --steps 5 --no-e2e --no-cpu --no-secondary --no-verify 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant $v', d['roofline']['kernel'], d['roofline']['stage_ms']['embed'], d['ms_per_step'])"; done
MMB_EMBED_PRESCALE=0 python bench.py --steps 5 --no-e2e --no-cpu --no-secondary --no-verify 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('prescale off', d['roofline']['kernel'], d['roofline']['stage_ms']['embed'], d['ms_per_step'])"
python multimodal-baselines_b200/sweep.py --limit 12 > gpurun_out/r02_sweep12_p1.json 2>/dev/null; cat gpurun_out/r02_sweep12_p1.json
for K in 2 3 4; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $K --master-addr 127.0.0.1 --master-port 2956$K multimodal-baselines_b200/sweep.py --limit 12 > gpurun_out/r02_sweep12_p$K.json 2>gpurun_out/r02_sweep12_p$K.err; cat gpurun_out/r02_sweep12_p$K.json; done
