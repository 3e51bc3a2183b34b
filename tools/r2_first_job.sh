# First GPU job of the next round (`gpurun --timeout 1200 -- 'bash tools/r2_first_job.sh'`, ~7 min): validates the experimental
# kernels that were written without hardware access and collects the A/B numbers the plan in DESIGN.md §7 asks for.
mkdir -p gpurun_out
# 1. the whole GPU suite, with the opt-in split-K tests included (no -x: see everything)
MAGGIE_B200_CONV_SPLITK=1 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_splitk.log 2>&1; echo "tests(+splitk) rc=$?"; tail -3 gpurun_out/r2_tests_splitk.log
# 2. per-layer conv table: default ring, 3-stage ring, split-K column
timeout 300 python tools/bench_conv_table.py --splitk > gpurun_out/r2_conv_table.txt 2>&1; echo "table rc=$?"
MAGGIE_B200_CONV_SMEM_KB=111 timeout 300 python tools/bench_conv_table.py > gpurun_out/r2_conv_table_smem111.txt 2>&1; echo "table(111) rc=$?"
tail -1 gpurun_out/r2_conv_table.txt; tail -1 gpurun_out/r2_conv_table_smem111.txt
# 3. whole-step A/B
for v in base smem111 splitk both; do
  case $v in base) e="";; smem111) e="MAGGIE_B200_CONV_SMEM_KB=111";; splitk) e="MAGGIE_B200_CONV_SPLITK=1";; both) e="MAGGIE_B200_CONV_SMEM_KB=111 MAGGIE_B200_CONV_SPLITK=1";; esac
  env $e timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_$v.json 2> gpurun_out/r2_bench_$v.err; echo "bench $v rc=$?"
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_bench_$v.json').read().strip().splitlines()[-1]); print('$v', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"
done
# 4. where the torch glue comes from (ATen device time by call site), eager step
STACKS=1 ROWS=60 timeout 300 python tools/profile_step.py > gpurun_out/r2_profile_stacks.txt 2>&1; echo "profile rc=$?"
# 5. launch list of one eager step
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_step.csv python tools/ncu_step.py > gpurun_out/r2_ncu_step.log 2>&1; echo "ncu rc=$?"
