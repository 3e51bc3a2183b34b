mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r1_bench_default.json 2> gpurun_out/r1_bench_default.err; echo "bench rc=$?"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_step.csv python tools/ncu_step.py > gpurun_out/s62_ncu_step.log 2>&1; echo "ncu step rc=$?"
ROWS=200 timeout 300 python tools/profile_step.py > gpurun_out/r1_profile_eager.txt 2>&1
