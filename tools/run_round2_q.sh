mkdir -p gpurun_out
NRW_AUX_STAGE=240 NRW_AUX_WARPS=12 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_optim.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
B="timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_bench16_$name.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench16_$name.json').read().strip().splitlines()[-1]);print('$name',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'])"; }
run default_1 X=1
run all12_1 NRW_AUX_STAGE=240 NRW_AUX_WARPS=12
run k456_12_1 NRW_AUX_STAGE=112 NRW_AUX_WARPS=12
run gate12_1 NRW_AUX_STAGE=16 NRW_AUX_WARPS=12
run none_1 NRW_AUX_STAGE=0
run default_2 X=1
run all12_2 NRW_AUX_STAGE=240 NRW_AUX_WARPS=12
rm -f /tmp/g.csv; NRW_AUX_STAGE=240 NRW_AUX_WARPS=12 NRW_GEMM_TIMING_DUMP=/tmp/g.csv $B > /dev/null 2>&1; python tools/gemm_table.py /tmp/g.csv 2>&1 | head -10
