mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_precision_policy.py 2>&1 | tail -4
B="timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
for i in 1 2; do
NRW_FUSED_HEAD=0 $B > gpurun_out/r2_bench15_nohead_$i.json 2>/dev/null
$B > gpurun_out/r2_bench15_head_$i.json 2>/dev/null
done
NRW_FUSED_HEAD=0 timeout 300 python bench.py --workload C5 --steps 3 --warmup 1 --no_torch_gpu_ref > gpurun_out/r2_bench15_c5_nohead.json 2>/dev/null
timeout 300 python bench.py --workload C5 --steps 3 --warmup 1 --no_torch_gpu_ref > gpurun_out/r2_bench15_c5_head.json 2>/dev/null
for f in nohead_1 head_1 nohead_2 head_2 c5_nohead c5_head; do python -c "
import json;d=json.loads(open('gpurun_out/r2_bench15_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'], d.get('loss'), d.get('sdf_min_max'))"; done
