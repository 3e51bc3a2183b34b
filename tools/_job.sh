# scratch job script for `gpurun -- 'bash tools/_job.sh'` (edited per experiment)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
MAGGIE_B200_BENCH_NO_THROTTLE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_nothrottle.json 2> gpurun_out/bench_nothrottle.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("bench", "bench_nothrottle"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["value"], 1), round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), d["clocks"])
PY
