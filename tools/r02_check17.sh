set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -2 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_final.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --no-verify > gpurun_out/plain_short.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --no-verify > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
