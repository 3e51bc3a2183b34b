set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s5_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s5_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; echo "bench rc=$?"
head -c 400 gpurun_out/s5_bench.json; tail -5 gpurun_out/s5_bench.err
ROWS=120 timeout 300 python tools/profile_step.py > gpurun_out/s5_profile_eager.txt 2>&1
GRAPHS=1 CPU_TABLE=1 ROWS=120 timeout 300 python tools/profile_step.py > gpurun_out/s5_profile_graphs.txt 2>&1
