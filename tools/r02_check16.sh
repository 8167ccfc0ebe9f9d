set -x
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 200 python tools/gram_probe.py 4000000 > gpurun_out/gram_probe5.jsonl 2>&1; cut -c1-160 gpurun_out/gram_probe5.jsonl
PROFILE_ITERS=1 python tools/profile_kernels.py > gpurun_out/plain_prof.log 2>&1 && PROFILE_ITERS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gram_tc_kernel" -c 1 -o gpurun_out/prof_r02e_gram -f python tools/profile_kernels.py > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log
