# fused SDF chain, second version (CTA-scope barrier ops, bias prefetch, one launch per query): check + timing + key ncu metrics
mkdir -p gpurun_out
for n in 1000 300001 2097152; do
NRW_SDF_FUSED=0 timeout -s KILL 120 python tools/fused_check.py /tmp/u_$n.pt $n 2>&1 | tail -1
NRW_SDF_FUSED=1 timeout -s KILL 120 python tools/fused_check.py /tmp/f_$n.pt $n 2>&1 | tail -2
timeout 60 python tools/fused_check.py cmp /tmp/u_$n.pt /tmp/f_$n.pt 2>&1 | tail -4
done
for f in 0 1; do
NRW_SDF_FUSED=$f timeout -s KILL 300 python bench.py --workload C5 --steps 3 --warmup 1 --no_torch_gpu_ref > gpurun_out/r2_fz3_c5_$f.json 2>gpurun_out/r2_fz3_c5_$f.err
python -c "
import json;d=json.loads(open('gpurun_out/r2_fz3_c5_$f.json').read().strip().splitlines()[-1]);print('c5 fused=$f',round(d['ms_per_step'],2),round(d['value']),d['clocks'], d.get('sdf_min_max'), round(d['roofline']['frac'],4), round(d['roofline']['mma_frac_of_peak'],4), d['gpu_launches'])" || tail -3 gpurun_out/r2_fz3_c5_$f.err
done
NRW_SDF_FUSED=1 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second --clock-control none --kernel-name-base demangled -k regex:"sdf_fused" -s 1 -c 1 python tools/fused_check.py /tmp/x.pt 2097152 2>&1 | grep -E "duration|tensor|dram|issue|per_second"
