# round-2 evidence sweep on one B200: tests, smoke, every bench workload + reference arm, launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2i_bench_c2_reference.json 2>/dev/null
timeout 600 python bench.py --workload C3 --steps 10 --warmup 3 > gpurun_out/r2i_bench_c3.json 2>/dev/null
timeout 600 python bench.py --workload C5 --steps 5 --warmup 1 > gpurun_out/r2i_bench_c5.json 2>/dev/null
timeout 600 python bench.py --workload C5 --impl reference --steps 2 --warmup 1 > gpurun_out/r2i_bench_c5_reference.json 2>/dev/null

for f in c2 c2_reference c3 c5 c5_reference; do python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2i_bench_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],2), 'ms', d.get('dtype'), 'e2e', round(d['e2e']['value'],1), 'frac', round(d.get('roofline',{}).get('frac',0),4), 'launches', d.get('gpu_launches'), (d.get('clocks') or {}).get('sm_mhz'), (d.get('clocks') or {}).get('reasons'))
    for k in ('reference_torch_gpu','cpu_baseline','other_precision_modes','octree_trace'):
        if k in d: print('   ',k, json.dumps(d[k])[:420])
except Exception as e: print('$f', 'ERR', e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/l.csv python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/l.csv 45 > gpurun_out/r2i_launches_step_mixed_summary.txt; head -14 gpurun_out/r2i_launches_step_mixed_summary.txt; rm -f gpurun_out/l.csv
