mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep -v "timed out" gpurun_out/smoke.log | tail -2 | cut -c1-300
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/tests_all.log 2>&1; echo "tests rc=$?"; grep -v "timed out" gpurun_out/tests_all.log | tail -4 | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --per-step > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_cfg2.json'))
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['per_step_ms'][:2])
P
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2.csv python tools/profile_step.py cfg2 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
