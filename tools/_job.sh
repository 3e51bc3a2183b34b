mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s28_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s28_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s28_bench.json 2> gpurun_out/s28_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/s28_bench.err
MAGGIE_B200_NO_PDL=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s28_bench_nopdl.json 2> gpurun_out/s28_bench_nopdl.err
