mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s48_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s48_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/s48_bench.json 2> gpurun_out/s48_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s48_bench.err
