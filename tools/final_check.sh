set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.json
python tools/bench_small.py > gpurun_out/bench_small.jsonl 2>/dev/null; tail -2 gpurun_out/bench_small.jsonl | cut -c1-300
python tools/bench_mmb.py --shape all --steps 100 > gpurun_out/bench_mmb.jsonl 2>gpurun_out/bm.err; grep -c metric gpurun_out/bench_mmb.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sif_embed|gram_tc|remove_pc|pcm_final|pcm_prep" -c 6 -o gpurun_out/prof_r1e -f python tools/profile_kernels.py > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
