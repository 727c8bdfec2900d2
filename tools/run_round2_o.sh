mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_optim.py tests/test_gpu_fine_sampling.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -5
B="timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
for i in 1 2; do
NRW_AUX_STAGE=0 $B > gpurun_out/r2_bench14_nostage_$i.json 2>/dev/null
$B > gpurun_out/r2_bench14_stage_$i.json 2>/dev/null
done
NRW_AUX_STAGE=224 $B > gpurun_out/r2_bench14_stage_nogatefwd.json 2>/dev/null
for f in nostage_1 stage_1 nostage_2 stage_2 stage_nogatefwd; do python -c "
import json;d=json.loads(open('gpurun_out/r2_bench14_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'],d['loss'])"; done
rm -f /tmp/g.csv; NRW_GEMM_TIMING_DUMP=/tmp/g.csv $B > /dev/null 2>&1; python tools/gemm_table.py /tmp/g.csv > gpurun_out/r2_gemm_table_mixed_fast7.txt 2>&1; head -12 gpurun_out/r2_gemm_table_mixed_fast6.txt
$B --precision bf16x3 > gpurun_out/r2_bench14_bf16x3.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench14_bf16x3.json').read().strip().splitlines()[-1]);print('bf16x3',round(d['ms_per_step'],2),round(d['value']))"
