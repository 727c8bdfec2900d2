# fused SDF chain: epilogue experiment (both TMEM loads in flight): parity tests, diff vs the per-layer chain, C5 timing, key ncu metrics
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sdf_grid.py tests/test_gpu_bitexact.py tests/test_gpu_parity2.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
NRW_SDF_FUSED=0 timeout -s KILL 120 python tools/fused_check.py /tmp/u.pt 2097152 2>&1 | tail -1
NRW_SDF_FUSED=1 timeout -s KILL 120 python tools/fused_check.py /tmp/f.pt 2097152 2>&1 | tail -1
timeout 60 python tools/fused_check.py cmp /tmp/u.pt /tmp/f.pt 2>&1 | head -1
for f in 1 1; do
NRW_SDF_FUSED=$f timeout -s KILL 300 python bench.py --workload C5 --steps 3 --warmup 1 --no_torch_gpu_ref > gpurun_out/r2_fz4_c5_$f.json 2>gpurun_out/r2_fz4_c5_$f.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_fz4_c5_$f.json').read().strip().splitlines()[-1]);print('c5 fused=$f',round(d['ms_per_step'],2),round(d['value']),d['clocks']['sm_mhz'], round(d['roofline']['mma_frac_of_peak'],4))" || tail -3 gpurun_out/r2_fz4_c5_$f.err
done
NRW_SDF_FUSED=1 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second --clock-control none --kernel-name-base demangled -k regex:"sdf_fused" -s 1 -c 1 python tools/fused_check.py /tmp/x.pt 2097152 2>&1 | grep -E "duration|tensor|issue|per_second"
