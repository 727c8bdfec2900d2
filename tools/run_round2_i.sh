B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
for i in 1 2; do
NRW_TC_DBG=4 $B > gpurun_out/r2_bench8_nopf_$i.json 2>/dev/null
$B > gpurun_out/r2_bench8_pf_$i.json 2>/dev/null
done
for f in nopf_1 pf_1 nopf_2 pf_2; do python -c "
import json;d=json.loads(open('gpurun_out/r2_bench8_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'])"; done
python -m pytest tests/test_gpu_dataio.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
