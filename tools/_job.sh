mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_model.py -x -q 2>&1 | tail -3
timeout 300 python tools/bench_sparse.py > gpurun_out/s31_sparse.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s31_bench.json 2> gpurun_out/s31_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s31_bench.err
