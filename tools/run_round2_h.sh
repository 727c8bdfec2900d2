mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -12 > gpurun_out/r2_tests7.log; tail -5 gpurun_out/r2_tests7.log
B="python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
NRW_EPI_FAST=0 NRW_BWD_GATE_PLANES=2 $B > gpurun_out/r2_bench7_generic.json 2>/dev/null
$B > gpurun_out/r2_bench7_fast.json 2>/dev/null
$B --precision bf16x3 > gpurun_out/r2_bench7_fast_bf16x3.json 2>/dev/null
for f in generic fast fast_bf16x3; do python -c "
import json;d=json.loads(open('gpurun_out/r2_bench7_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['ms_per_step'],2),round(d['value']),d['forward_slots'],round(d['roofline']['frac'],4),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'],d['loss'])"; done
rm -f /tmp/g.csv; NRW_GEMM_TIMING_DUMP=/tmp/g.csv $B > /dev/null 2>&1; python tools/gemm_table.py /tmp/g.csv > gpurun_out/r2_gemm_table_mixed_fast3.txt 2>&1; head -14 gpurun_out/r2_gemm_table_mixed_fast3.txt
