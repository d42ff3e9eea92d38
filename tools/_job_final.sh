mkdir -p gpurun_out
B="--config cfg3 --steps 10 --warmup 3 --both-precisions 0 --strong-cfg4 0 --no-cpu-baseline --per-step"
for v in default gx0 dg0; do
  case $v in default) E="";; gx0) E="CTCB200_OVERLAP_GX=0";; dg0) E="CTCB200_OVERLAP_DG=0";; esac
  env $E timeout 200 python bench.py $B > gpurun_out/bench_cfg3_$v.json 2> gpurun_out/bench_cfg3_$v.err; echo "cfg3 $v rc=$?"
  python - <<P
import json
d=json.load(open('gpurun_out/bench_cfg3_$v.json'))
print('$v', round(d['ms_per_step'],3), round(32e3/d['e2e']['value'],3), d['per_step_ms'])
P
done
