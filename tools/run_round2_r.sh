# 256 x 512 pair tiles for the one-plane weight-gradient GEMMs: parity tests with the tile on, then a same-box A/B
mkdir -p gpurun_out
timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_optim.py tests/test_gpu_engine_state.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
B="timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_bench17_$name.json 2>gpurun_out/r2_bench17_$name.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench17_$name.json').read().strip().splitlines()[-1]);print('$name',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'], d.get('loss'))" || tail -3 gpurun_out/r2_bench17_$name.err; }
run narrow_1 NRW_DW_WIDE=0
run wide_1 NRW_DW_WIDE=1
run wide2_1 NRW_DW_WIDE=2
run narrow_2 NRW_DW_WIDE=0
run wide_2 NRW_DW_WIDE=1
rm -f /tmp/g.csv; NRW_DW_WIDE=1 NRW_GEMM_TIMING_DUMP=/tmp/g.csv $B > /dev/null 2>&1; python tools/gemm_table.py /tmp/g.csv > gpurun_out/r2_gemm_table_wide.txt 2>&1; head -30 gpurun_out/r2_gemm_table_wide.txt
rm -f /tmp/g.csv; NRW_DW_WIDE=0 NRW_GEMM_TIMING_DUMP=/tmp/g.csv $B > /dev/null 2>&1; python tools/gemm_table.py /tmp/g.csv > gpurun_out/r2_gemm_table_narrow.txt 2>&1; head -30 gpurun_out/r2_gemm_table_narrow.txt
