mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_step.csv python tools/ncu_step.py > gpurun_out/s37_ncu_step.log 2>&1; echo "ncu step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_halo|sparse_conv_persistent|sparse_wgrad_persistent|conv_tcgen05|wgrad_tcgen05" -s 5 -c 5 -f -o gpurun_out/r1_full python tools/ncu_kernels.py > gpurun_out/s37_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/s37_ncu_full.log
