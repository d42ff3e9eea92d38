mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ctc" > gpurun_out/tests_ctc.log 2>&1; echo "tests rc=$?"; grep -v "timed out" gpurun_out/tests_ctc.log | tail -3 | cut -c1-300
timeout 200 python tools/ctc_sweep_bench.py gpurun_out/ctc_sweep.json > gpurun_out/ctc_sweep.log 2>&1; echo "ctc_sweep rc=$?"; python - <<'P'
import json
for r in json.load(open('gpurun_out/ctc_sweep.json'))['rows']:
    print(r['N'], r['form'], round(r['frac_of_hbm_peak'],4), {k:(round(v['fwd'],3),round(v['bwd'],3)) for k,v in r['both_forms_ms'].items()})
P
timeout 300 python bench.py --steps 10 --warmup 3 --both-precisions 0 --strong-cfg4 0 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k:v for k,v in d['rooflines_other']['kernel_ms_per_step'].items() if 'ctc' in k})
P
