set -x
mkdir -p gpurun_out
rm -f gpurun_out/mmb_test_errors.tsv
MMB_TEST_ERRLOG=gpurun_out/mmb_test_errors.tsv python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu12.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu12.log | cut -c1-250
python multimodal-baselines_b200/sweep.py --limit 16 --regressor-batch 1 > gpurun_out/r02_sweep16_rb1.json 2>/dev/null; tail -1 gpurun_out/r02_sweep16_rb1.json | cut -c1-600
python multimodal-baselines_b200/sweep.py --limit 16 --regressor-batch 8 --out gpurun_out/r02_sweep16_rb8.jsonl > gpurun_out/r02_sweep16_rb8.json 2>gpurun_out/r02_sweep16_rb8.err; tail -1 gpurun_out/r02_sweep16_rb8.json | cut -c1-600; tail -3 gpurun_out/r02_sweep16_rb8.err
python multimodal-baselines_b200/sweep.py --limit 16 --regressor-batch 16 > gpurun_out/r02_sweep16_rb16.json 2>/dev/null; tail -1 gpurun_out/r02_sweep16_rb16.json | cut -c1-600
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2>/dev/null; cut -c1-400 gpurun_out/r02_bench_ref.json
