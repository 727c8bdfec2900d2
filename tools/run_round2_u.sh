# fused forward-only SDF chain: sampler / parity tests with it on, C5 and C2 same-box A/B, ncu --set full of the kernel
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
NRW_SDF_FUSED=1 timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py tests/test_gpu_bitexact.py tests/test_gpu_sdf_grid.py tests/test_gpu_fine_sampling.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
for i in 1 2; do
for f in 0 1; do
NRW_SDF_FUSED=$f timeout -s KILL 300 python bench.py --workload C5 --steps 3 --warmup 1 --no_torch_gpu_ref > gpurun_out/r2_fz_c5_${f}_$i.json 2>gpurun_out/r2_fz_c5_${f}_$i.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_fz_c5_${f}_$i.json').read().strip().splitlines()[-1]);print('c5 fused=$f',round(d['ms_per_step'],2),round(d['value']),d['clocks'], d.get('sdf_min_max'), round(d['roofline']['frac'],4), round(d['roofline']['mma_frac_of_peak'],4))" || tail -3 gpurun_out/r2_fz_c5_${f}_$i.err
done; done
B="timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
for i in 1 2; do for f in 0 1; do
NRW_SDF_FUSED=$f $B > gpurun_out/r2_fz_c2_${f}_$i.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2_fz_c2_${f}_$i.json').read().strip().splitlines()[-1]);print('c2 fused=$f',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'], d.get('loss'), round(d['roofline']['frac'],4))"
done; done
NRW_SDF_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"sdf_fused" -s 2 -c 1 -o gpurun_out/r2h_sdf_fused python tools/fused_check.py /tmp/x.pt 2097152 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
