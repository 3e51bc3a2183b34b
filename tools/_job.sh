mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_model.py tests/test_gpu_loss.py -x -q 2>&1 | tail -2
timeout 300 python tools/bench_sparse.py > gpurun_out/s25_sparse.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s25_bench.json 2> gpurun_out/s25_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s25_bench.err
