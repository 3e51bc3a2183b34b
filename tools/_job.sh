mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loss.py tests/test_gpu_model.py -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/s40_bench.json 2> gpurun_out/s40_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s40_bench.err
