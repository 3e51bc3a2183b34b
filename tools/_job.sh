set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_weights.py -x -q > gpurun_out/s3_wtests.log 2>&1; echo "wtests rc=$?"; tail -15 gpurun_out/s3_wtests.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s3_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; echo "bench rc=$?"
head -c 600 gpurun_out/s3_bench.json; tail -5 gpurun_out/s3_bench.err
ROWS=90 timeout 300 python tools/profile_step.py > gpurun_out/s3_profile_eager.txt 2>&1
