mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6
B="timeout 300 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_torch_gpu_ref --no_other_modes"
$B > gpurun_out/r2_bench12.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r2_bench12.json').read().strip().splitlines()[-1]);print('mixed',round(d['ms_per_step'],2),round(d['value']),round(d['roofline']['kernel_ms_per_step'],1),d['clocks']['sm_mhz'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/l.csv python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/l.csv 40 > gpurun_out/r2_launches_step_mixed3_summary.txt; sed -n 10,36p gpurun_out/r2_launches_step_mixed3_summary.txt; rm -f gpurun_out/l.csv
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"tc2_kernel.*6>" -s 8 -c 1 -o gpurun_out/r2_k6_reverse python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"tc2_kernel.*4>" -s 8 -c 1 -o gpurun_out/r2_k4_gatefwd python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k regex:"tc2_kernel.*0, .*1>" -s 60 -c 1 -o gpurun_out/r2_k1_fwd python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"composite" -c 2 -o gpurun_out/r2_composite2 python tools/prof_step.py --precision mixed --chunk_rows 262144 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
