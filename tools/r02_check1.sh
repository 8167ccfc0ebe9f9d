set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -4 gpurun_out/r02_smoke.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --no-verify > gpurun_out/plain_short.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-secondary --no-verify > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log | cut -c1-200
python tools/profile_kernels.py > gpurun_out/plain_prof.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sif_embed|gram_tc|remove_pc|pcm_final|pcm_prep" -c 6 -o gpurun_out/prof_r02 -f python tools/profile_kernels.py > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
MMB_BENCH_IDS=uniform PROFILE_ITERS=1 python tools/profile_kernels.py > gpurun_out/plain_prof_u.log 2>&1 && MMB_BENCH_IDS=uniform PROFILE_ITERS=1 timeout 400 ncu --set full --clock-control none -k regex:"sif_embed" -c 1 -o gpurun_out/prof_r02_uniform -f python tools/profile_kernels.py > gpurun_out/ncu_u.log 2>&1; tail -2 gpurun_out/ncu_u.log
