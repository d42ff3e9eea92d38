mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep -v "timed out" gpurun_out/smoke.log | tail -2 | cut -c1-300
ncu env 2>/dev/null | grep -i -E "inject|nsight|profiler|nv_" > gpurun_out/ncu_env.txt; cat gpurun_out/ncu_env.txt | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/smoke_ncu.csv python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_ncu.log 2>&1; echo "smoke under ncu rc=$?"; grep -v "timed out" gpurun_out/smoke_ncu.log | tail -2 | cut -c1-300; grep -c ctcb200 gpurun_out/smoke_ncu.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2.csv python tools/profile_step.py cfg2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"; grep -v "timed out" gpurun_out/ncu_launches.log | tail -2 | cut -c1-200
CUDA_LAUNCH_BLOCKING=1 timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_blocking.log 2>&1; echo "smoke CUDA_LAUNCH_BLOCKING rc=$?"
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/tests_all.log 2>&1; echo "tests rc=$?"; grep -v "timed out" gpurun_out/tests_all.log | tail -4 | cut -c1-300
