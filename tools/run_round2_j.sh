mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider -s 2>&1 | grep -E "grad_cosine|passed|failed|FAILED|Error|assert " | head -30 > gpurun_out/r2_tests8.log; cat gpurun_out/r2_tests8.log
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
for i in 1 2; do
NRW_AUX_BF16=0 $B > gpurun_out/r2_bench9_f32aux_$i.json 2>/dev/null
$B > gpurun_out/r2_bench9_bf16aux_$i.json 2>/dev/null
done
for f in f32aux_1 bf16aux_1 f32aux_2 bf16aux_2; do python -c "
import json;d=json.loads(open('gpurun_out/r2_bench9_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['kernel_ms_per_step'],1),round(d['roofline']['algorithmic_hbm_gb_per_step_in_kernel']),d['clocks']['sm_mhz'],d['forward_slots'])"; done
rm -f /tmp/g.csv; NRW_GEMM_TIMING_DUMP=/tmp/g.csv $B > /dev/null 2>&1; python tools/gemm_table.py /tmp/g.csv > gpurun_out/r2_gemm_table_mixed_fast4.txt 2>&1; head -14 gpurun_out/r2_gemm_table_mixed_fast4.txt
