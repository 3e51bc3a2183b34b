set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -x -q > gpurun_out/s8_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/s8_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err; echo "bench rc=$?"
head -c 300 gpurun_out/s8_bench.json; tail -5 gpurun_out/s8_bench.err
timeout 300 python tools/cpu_profile.py > gpurun_out/s8_cpu_profile.txt 2>&1
