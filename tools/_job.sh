set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_tests.log 2>&1; echo "tests rc=$?" 
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/s2_bench.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches.csv python tools/ncu_step.py > gpurun_out/s2_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/s2_launches.csv
ROWS=70 timeout 300 python tools/profile_step.py > gpurun_out/s2_profile_eager.txt 2>&1
tail -3 gpurun_out/s2_tests.log
