mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rows.py tests/test_gpu_model.py tests/test_gpu_sparse.py tests/test_gpu_attention.py -x -q > gpurun_out/s12_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/s12_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s12_bench.json 2> gpurun_out/s12_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s12_bench.err
GRAPHS=1 ROWS=150 timeout 300 python tools/profile_step.py > gpurun_out/s12_profile_graphs.txt 2>&1
ROWS=150 timeout 300 python tools/profile_step.py > gpurun_out/s12_profile_eager.txt 2>&1
