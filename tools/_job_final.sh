mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2.csv python tools/profile_step.py cfg2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
