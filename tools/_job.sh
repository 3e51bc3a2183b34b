mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_conv_bn.py tests/test_gpu_sparse.py -x -q > gpurun_out/s11_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s11_tests.log
timeout 600 python tools/bench_conv_table.py > gpurun_out/s11_conv_table.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err; echo "bench rc=$?"
