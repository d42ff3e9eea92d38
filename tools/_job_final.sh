mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep -v "timed out" gpurun_out/smoke.log | tail -3 | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; head -c 700 gpurun_out/bench_cfg2.json; grep -v "timed out" gpurun_out/bench_cfg2.err | tail -3 | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2.csv python tools/profile_step.py cfg2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streamed" > gpurun_out/tests_stream.log 2>&1; echo "tests streamed (fresh process) rc=$?"; grep -v "timed out" gpurun_out/tests_stream.log | tail -3 | cut -c1-300
