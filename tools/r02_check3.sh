set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "ragged or device_masks or sweep_grid or embed_shapes or golden_pipeline or nonfinite" > gpurun_out/r02_pytest_new.log 2>&1; tail -8 gpurun_out/r02_pytest_new.log
python multimodal-baselines_b200/sweep.py --limit 8 --out gpurun_out/r02_sweep8.jsonl > gpurun_out/r02_sweep8_summary.json 2> gpurun_out/r02_sweep8.err; cat gpurun_out/r02_sweep8_summary.json; head -c 1500 gpurun_out/r02_sweep8.jsonl
python tools/bench_mmb.py --shape mosi --steps 30 --no-cpu --only-eager-ours > gpurun_out/plain_mmb.log 2>&1 && timeout 600 ncu --set full --clock-control none -k regex:"heads_|word_|gauss_|splitk|scale_multi|gather_multi|row_inv" -s 200 -c 40 -o gpurun_out/prof_r02_mmb -f python tools/bench_mmb.py --shape mosi --steps 30 --no-cpu --only-eager-ours > gpurun_out/ncu_mmb.log 2>&1; tail -2 gpurun_out/ncu_mmb.log
