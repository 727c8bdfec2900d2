# programmatic dependent launch of the tcgen05 GEMMs (NRW_PDL) and 128-wide layers on the CTA-pair kernel (NRW_PAIR_MIN_N=128):
# parity tests with both on, then a same-box A/B
mkdir -p gpurun_out
NRW_PDL=1 NRW_PAIR_MIN_N=128 timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_bitexact.py tests/test_gpu_engine_state.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
B="timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
run() { name=$1; shift; env "$@" $B > gpurun_out/r2_bench18_$name.json 2>gpurun_out/r2_bench18_$name.err; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench18_$name.json').read().strip().splitlines()[-1]);print('$name',round(d['ms_per_step'],2),round(d['value']),round(d['e2e']['ms_per_step'],2),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'], d.get('loss'), d.get('stages'))" || tail -3 gpurun_out/r2_bench18_$name.err; }
run off_1 NRW_PDL=0
run pdl_1 NRW_PDL=1
run pair128_1 NRW_PAIR_MIN_N=128
run both_1 NRW_PDL=1 NRW_PAIR_MIN_N=128
run off_2 NRW_PDL=0
run pdl_2 NRW_PDL=1
run pair128_2 NRW_PAIR_MIN_N=128
run both_2 NRW_PDL=1 NRW_PAIR_MIN_N=128
rm -f /tmp/g.csv; NRW_PAIR_MIN_N=128 NRW_GEMM_TIMING_DUMP=/tmp/g.csv $B > /dev/null 2>&1; python tools/gemm_table.py /tmp/g.csv > gpurun_out/r2_gemm_table_pair128.txt 2>&1; grep -E " (128|64) +(128|64|256|640) " gpurun_out/r2_gemm_table_pair128.txt | head
NRW_PDL=0 timeout 300 python bench.py --workload C5 --steps 3 --warmup 1 --no_torch_gpu_ref > gpurun_out/r2_bench18_c5_off.json 2>/dev/null
NRW_PDL=1 timeout 300 python bench.py --workload C5 --steps 3 --warmup 1 --no_torch_gpu_ref > gpurun_out/r2_bench18_c5_on.json 2>/dev/null
for f in c5_off c5_on; do python -c "
import json;d=json.loads(open('gpurun_out/r2_bench18_$f.json').read().strip().splitlines()[-1]);print('$f',round(d['ms_per_step'],2),round(d['value']),d['clocks']['sm_mhz'], d.get('sdf_min_max'))"; done
