mkdir -p gpurun_out
timeout -s KILL 100 python tools/rec_ab_head.py 800 32 512 > gpurun_out/rec_ab_head.json 2> gpurun_out/rec_ab_head.err; echo "rec rc=$?"; cat gpurun_out/rec_ab_head.json; tail -2 gpurun_out/rec_ab_head.err | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; grep -v "timed out" gpurun_out/smoke.log | tail -1 | cut -c1-300
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/tests_all.log 2>&1; echo "tests rc=$?"; grep -v "timed out" gpurun_out/tests_all.log | tail -3 | cut -c1-300
timeout 400 python bench.py --steps 10 --warmup 3 --per-step > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_cfg2.json'))
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['per_step_ms'][:2], d['roofline']['avg_launch_ms'], d['other_precision']['ms_per_step'], d['strong_cfg4']['ms_per_step'])
P
