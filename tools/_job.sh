# scratch job script for `gpurun -- 'bash tools/_job.sh'` (edited per experiment)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 200 python tools/bench_optim.py > gpurun_out/optim.log 2>&1; tail -3 gpurun_out/optim.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1; echo "ncu rc=$?"
